"""CPU checks of the reference-comparison harness (tests/ref_harness.py) and of the staged reference:
 * oracle/_ref holds the reference files bit-identical to the digests pinned in oracle/ref_manifest.json;
 * the harness, run with the numpy oracle as the candidate, passes against the LIVE reference module on CPU (this also
   pins the oracle against the executed reference beyond the committed golden vectors) -- dense, permuted, multi-step,
   row-chunked reference;
 * the harness rejects a wrong index, a perturbed EMA buffer and a wrong quantize value (it can fail).
"""
import numpy as np
import pytest
import torch

import ref_harness as H
from oracle import reference_module
from oracle.quantize_oracle import QuantizeOracle


@pytest.fixture(scope="module")
def ref():
    try:
        return reference_module.load("vqvae")
    except reference_module.ReferenceUnavailable as exc:
        pytest.skip(f"reference not staged: {exc}")


class OracleModule:
    """The numpy oracle behind the module surface the harness drives (buffers as torch tensors, in-place updates)."""

    def __init__(self, ref_q):
        self.dim, self.n_embed, self.decay, self.eps = ref_q.dim, ref_q.n_embed, ref_q.decay, ref_q.eps
        self.training = ref_q.training
        self.embed = ref_q.embed.detach().clone()
        self.cluster_size = ref_q.cluster_size.detach().clone()
        self.embed_avg = ref_q.embed_avg.detach().clone()
        self.fault = None

    def __call__(self, x):
        o = QuantizeOracle(self.dim, self.n_embed, self.decay, self.eps, embed=self.embed.numpy())
        o.load(self.embed.numpy(), self.cluster_size.numpy(), self.embed_avg.numpy())
        o.training = self.training
        q, d, i = o.forward(np.ascontiguousarray(x.numpy()))
        if self.fault == "index":
            i = i.copy(); i.reshape(-1)[3] = (i.reshape(-1)[3] + 1) % self.n_embed
        if self.fault == "quantize":
            q = q.copy(); q.reshape(-1)[5] *= 1.0 + 1e-4
        if self.training:
            self.embed.copy_(torch.from_numpy(o.embed)); self.cluster_size.copy_(torch.from_numpy(o.cluster_size))
            self.embed_avg.copy_(torch.from_numpy(o.embed_avg))
            if self.fault == "ema":
                self.embed_avg[7, 11] *= 1.0 + 1e-4
        qt = torch.empty_strided(x.shape, x.stride(), dtype=torch.float32)
        qt.copy_(torch.from_numpy(q))
        return qt, torch.tensor(d), torch.from_numpy(i)


def test_staged_reference_matches_pinned_digests(ref):
    files = reference_module.verify()
    assert "vqvae.py" in files and "distributed/distributed.py" in files
    assert hasattr(ref, "Quantize") and hasattr(ref, "VQVAE")


@pytest.mark.parametrize("permuted", [False, True])
def test_harness_passes_oracle_vs_live_reference(ref, permuted):
    torch.manual_seed(0)
    r = ref.Quantize(64, 512).train()
    o = OracleModule(r)
    for step, kind in enumerate(["randn", "randn", "clustered"]):
        x = H.make_inputs(kind, (2, 16, 16, 64), r.embed.detach(), 10 + step, "cpu", permuted)
        e = H.compare_step(f"cpu-{step}", r, o, x)
        assert e["quantize"] <= 1e-6 and e["embed_avg"] <= H.TOL
    r.eval(); o.training = False
    x = H.make_inputs("randn", (3, 8, 8, 64), r.embed.detach(), 99, "cpu", permuted)
    H.compare_step("cpu-eval", r, o, x)


def test_harness_chunked_reference_equals_plain_reference(ref):
    torch.manual_seed(1)
    a = ref.Quantize(32, 128).train()
    b = ref.Quantize(32, 128).train()
    b.load_state_dict(a.state_dict())
    x = torch.randn(1000, 32, generator=torch.Generator().manual_seed(5))
    qa, da, ia = a(x)
    qb, db, ib = H.chunked_reference_forward(b, x, 96)
    assert torch.equal(ia, ib) and torch.equal(qa, qb)
    assert abs(float(da) - float(db)) <= 1e-6 * float(da)
    assert torch.allclose(a.cluster_size, b.cluster_size, rtol=1e-6, atol=0)
    assert torch.allclose(a.embed_avg, b.embed_avg, rtol=0, atol=1e-5 * float(a.embed_avg.abs().max()))
    assert b.training


@pytest.mark.parametrize("fault", ["index", "quantize", "ema"])
def test_harness_can_fail(ref, fault):
    torch.manual_seed(2)
    r = ref.Quantize(64, 512).train()
    o = OracleModule(r)
    o.fault = fault
    x = H.make_inputs("randn", (512, 64), r.embed.detach(), 3, "cpu")
    with pytest.raises(AssertionError):
        H.compare_step(f"fault-{fault}", r, o, x)


def test_harness_tolerates_only_true_near_ties(ref):
    """A row exactly between two codes may resolve either way; a row clearly closer to one code may not."""
    torch.manual_seed(3)
    r = ref.Quantize(64, 512).eval()
    e = r.embed.detach()
    x = torch.randn(64, 64)
    x[0] = 0.5 * (e[:, 20] + e[:, 21])
    _, _, ri = r(x)
    other = ri.clone()
    other[0] = 21 if int(ri[0]) == 20 else 20
    n_differ, n_bad, _, codes = H.index_mismatches(x, e, other, ri)
    assert (n_differ, n_bad) == (1, 0) and set(codes.tolist()) == {20, 21}
    other[5] = (int(ri[5]) + 1) % 512
    assert H.index_mismatches(x, e, other, ri)[1] == 1
