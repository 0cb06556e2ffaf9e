"""NVLink flag round-trip latency between two GPUs over symmetric (peer-mapped) memory: the floor of any exchange step.
Run under torchrun with exactly 2 ranks (one per GPU)."""
import ctypes as C
import os
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from vq_vae_2_pytorch_b200 import _native  # noqa: E402

rank = int(os.environ.get("RANK", 0))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
assert dist.get_world_size() == 2
import torch.distributed._symmetric_memory as symm  # noqa: E402
lib = _native.load()
buf = symm.empty(64, dtype=torch.int32, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
ptrs = [int(p) for p in hdl.buffer_ptrs]
ns = torch.zeros(1, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
for fence in (0, 1):
    for rep in range(2):
        buf.zero_()
        torch.cuda.synchronize()
        hdl.barrier()
        iters = 2000
        _native.check(lib.vqb200_debug_pingpong(C.c_void_p(ptrs[rank]), C.c_void_p(ptrs[1 - rank]), iters, 1 if rank == 0 else 0, fence,
                                                C.c_void_p(ns.data_ptr()), st), "pingpong")
        torch.cuda.synchronize()
        if rank == 0 and rep == 1:
            print(f"flag round trip over peer memory ({'fence.sys + ' if fence else ''}st.release.sys -> ld.acquire.sys spin): "
                  f"{int(ns.item()) / iters / 1e3:.2f} us per round trip")
dist.destroy_process_group()
