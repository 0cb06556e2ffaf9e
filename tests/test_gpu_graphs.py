"""The forward is CUDA-graph capturable (reference trainers at B = 8 .. 32 per GPU are launch-bound: train_vqvae.py:85-100):
no host synchronisation, no pinned read-back and no event query inside a captured forward; replays reproduce the eager
results, training replays advance the EMA exactly like eager calls."""
import pytest
import torch

import vq_vae_2_pytorch_b200 as vq

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("layout", ["dense", "nchw"])
def test_eval_forward_replays_from_a_cuda_graph(layout):
    torch.manual_seed(0)
    q = vq.Quantize(64, 512).to(DEV).eval()
    x = torch.randn(8, 64, 32, 32, device=DEV).permute(0, 2, 3, 1) if layout == "nchw" else torch.randn(8, 32, 32, 64, device=DEV)
    static_x = x.clone(memory_format=torch.preserve_format)
    for _ in range(3):
        q(static_x)                                   # warm-up: workspaces, function attributes, tensor maps
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        quant, diff, ind = q(static_x)
    for seed in (1, 2, 3):
        new = torch.randn(x.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(seed))
        static_x.copy_(new)
        g.replay()
        torch.cuda.synchronize()
        eq, ed, ei = q(new.clone(memory_format=torch.preserve_format) if layout == "dense" else static_x.clone(memory_format=torch.preserve_format))
        assert torch.equal(ind, ei) and torch.equal(quant, eq) and abs(float(diff) - float(ed)) <= 1e-6 * float(ed)


def test_training_steps_replay_from_a_cuda_graph():
    """Three training forwards captured in one graph (each sees the previous one's EMA update); replaying it equals the
    same calls issued eagerly on a twin module."""
    torch.manual_seed(1)
    a = vq.Quantize(64, 512).to(DEV).train()
    b = vq.Quantize(64, 512).to(DEV).train()
    b.load_state_dict(a.state_dict())
    xs = [torch.randn(4, 32, 32, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(10 + i)) for i in range(3)]
    warm = vq.Quantize(64, 512).to(DEV).train()
    for x in xs:
        warm(x)                                       # process-wide one-time set-up outside the capture
    torch.cuda.synchronize()
    ws = a._workspace(torch.device(DEV), xs[0].shape[0] * 32 * 32)     # workspaces exist before the capture
    assert ws is not None
    state0 = {k: v.clone() for k, v in a.state_dict().items()}
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [a(x) for x in xs]
    a.load_state_dict(state0)                         # capture does not execute: start the replay from the initial state
    g.replay()
    torch.cuda.synchronize()
    for step, (x, (quant, diff, ind)) in enumerate(zip(xs, outs)):
        eq, ed, ei = b(x)
        assert torch.equal(ind, ei)
        # step 0 starts from identical buffers: identical bits; later steps gather from a codebook that went through the
        # statistics kernel, whose summation order is not reproducible between two runs (like the reference's GEMM)
        if step == 0:
            assert torch.equal(quant, eq)
        else:
            assert torch.allclose(quant, eq, rtol=1e-5, atol=1e-6)
    for name in ("cluster_size", "embed_avg", "embed"):
        ga, gb = getattr(a, name), getattr(b, name)
        assert torch.allclose(ga, gb, rtol=1e-5, atol=1e-6 * float(gb.abs().max())), name
