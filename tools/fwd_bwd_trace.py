import sys, torch
sys.path.insert(0, ".")
import vq_vae_2_pytorch_b200 as vq
dev="cuda:0"; D,K,N=64,512,128*64*64
torch.manual_seed(0)
q=vq.Quantize(D,K).to(dev).train()
pick=torch.randint(0,K,(N,),device=dev)
x=(q.embed.t()[pick]+0.1*torch.randn(N,D,device=dev)).reshape(128,64,64,D).contiguous()
if len(sys.argv)>1 and sys.argv[1]=='nchw': x=x.permute(0,3,1,2).contiguous().permute(0,2,3,1)
x=x.requires_grad_(True)
def step():
    quant,diff,ind=q(x)
    (quant.sum()+0.25*diff).backward()
    x.grad=None
for _ in range(5): step()
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): step()
b.record(); torch.cuda.synchronize()
print("fwd+bwd (incl. torch sum/backward glue) us/step", a.elapsed_time(b)/20*1e3)
from torch.profiler import profile, ProfilerActivity
from collections import defaultdict
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
tot=defaultdict(float)
for ev in prof.events():
    if ev.device_type==torch.autograd.DeviceType.CUDA: tot[ev.name[:60]]+=ev.device_time
for n,t in sorted(tot.items(), key=lambda kv:-kv[1])[:12]: print(f"{t/5:9.1f} us  {n}")
