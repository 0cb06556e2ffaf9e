"""Index egress (extract_code.py:23-33): narrowed codes over PCIe, pinned ring, push order."""
import numpy as np
import pytest
import torch

import vq_vae_2_pytorch_b200 as vq
from vq_vae_2_pytorch_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("K,nbytes", [(512, 2), (65536, 2), (70000, 4)])
def test_pack_unpack_round_trip(K, nbytes):
    lib = _native.load()
    g = torch.Generator(device=DEV).manual_seed(K)
    for n in (0, 1, 3, 4, 1021, 128 * 64 * 64 + 5):
        ind = torch.randint(0, K, (n,), device=DEV, generator=g)
        if n:
            ind[-1] = K - 1
        out = torch.empty(max(n, 1), dtype=torch.int16 if nbytes == 2 else torch.int32, device=DEV)
        status = torch.full((1,), 7, dtype=torch.int32, device=DEV)
        st = _native.C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _native.check(lib.vqb200_pack_indices(_native.ptr(ind), n, K, nbytes, _native.ptr(out), _native.ptr(status), st), "pack")
        back = vq.unpack_codes(out[:n]) if n else ind
        assert torch.equal(back, ind)
        if n:
            assert int(status.item()) == 0
    assert lib.vqb200_pack_indices(_native.ptr(ind), 8, 70000, 2, _native.ptr(out), None, None) == -2      # 16 bits too few
    assert lib.vqb200_pack_indices(_native.ptr(ind), 8, 512, 3, _native.ptr(out), None, None) == -1


def test_out_of_range_index_is_reported():
    eg = vq.CodeEgress(512)
    bad = torch.tensor([1, 2, 600, 3], device=DEV)
    eg.push(bad)
    with pytest.raises(RuntimeError, match="outside"):
        eg.pop()


def test_code_egress_matches_reference_copy_and_order():
    """What extract_code.py stores per batch: id_t [B,32,32] and id_b [B,64,64] as numpy int64."""
    torch.manual_seed(0)
    qt, qb = vq.Quantize(64, 512).to(DEV).eval(), vq.Quantize(64, 512).to(DEV).eval()
    eg = vq.CodeEgress(512, depth=2)
    want = []
    for step in range(5):
        B = 3 + step
        id_t = qt.assign(torch.randn(B, 32, 32, 64, device=DEV))
        id_b = qb.assign(torch.randn(B, 64, 64, 64, device=DEV))
        nbytes = eg.push(id_t, id_b, tag=step)
        assert nbytes == 2 * (id_t.numel() + id_b.numel())          # 4x fewer than the reference's int64 copy
        want.append((id_t.cpu().numpy(), id_b.cpu().numpy()))      # the reference's own path
        if step >= 1:                                               # consumer one batch behind the encoder
            tag, (top, bottom) = eg.pop()
            assert tag == step - 1
            assert top.dtype == np.int64 and top.shape == want[tag][0].shape
            assert np.array_equal(top, want[tag][0]) and np.array_equal(bottom, want[tag][1])
    rest = list(eg.drain())
    assert [t for t, _ in rest] == [4] and np.array_equal(rest[0][1][1], want[4][1])
    narrow = vq.CodeEgress(512, widen=False)
    narrow.push(id_b)
    _, (nb,) = narrow.pop()
    assert nb.dtype == np.uint16 and np.array_equal(nb.astype(np.int64), want[4][1])
