"""Pin the CPU oracle (oracle/quantize_oracle.py) against outputs of the reference module.

The fixtures were produced by tests/golden/make_golden.py, which executes the reference
`Quantize` (/root/reference/vqvae.py:28-78) on seeded inputs.  CPU-only.
"""
import numpy as np
import pytest

from helpers import REL_TOL, check_outputs_np, elem_rel_err, golden_names, load_golden, rel_err
from oracle.quantize_oracle import QuantizeOracle, tie_tolerant_index_mismatches


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_outputs(name):
    g = load_golden(name)
    o = QuantizeOracle(int(g["dim"]), int(g["n_embed"]), float(g["decay"]), float(g["eps"]), embed=g["embed0"])
    o.load(g["embed0"], g["cluster_size0"], g["embed_avg0"])
    o.training = bool(g["train"])
    for s in range(int(g["steps"])):
        x = np.ascontiguousarray(g[f"x{s}"])
        embed_before = o.embed.copy()
        before = o.state()
        q, diff, ind = o.forward(x)
        assert ind.dtype == np.int64 and ind.shape == x.shape[:-1]
        train = bool(g["train"])
        check_outputs_np(f"{name} step {s}", x, before, (q, diff, ind), (g[f"quantize{s}"], g[f"diff{s}"], g[f"ind{s}"]),
                         (o.cluster_size, o.embed_avg, o.embed) if train else None,
                         (g[f"cluster_size{s + 1}"], g[f"embed_avg{s + 1}"], g[f"embed{s + 1}"]) if train else None,
                         train, float(g["decay"]), float(g["eps"]))
        # keep both trajectories on the reference's state so later steps stay comparable
        o.load(g[f"embed{s + 1}"], g[f"cluster_size{s + 1}"], g[f"embed_avg{s + 1}"])
        if f"xgrad{s}" in g:
            codes = embed_before.T[g[f"ind{s}"]]
            gx = o.backward(x, codes, g[f"gq{s}"], float(g[f"gd{s}"]))
            assert elem_rel_err(gx, g[f"xgrad{s}"], floor=1e-6 * float(np.abs(g[f"xgrad{s}"]).max())) <= REL_TOL


def test_oracle_chunked_equals_unchunked():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((777, 64)).astype(np.float32)
    a = QuantizeOracle(64, 128, seed=1)
    b = QuantizeOracle(64, 128, seed=1)
    qa, da, ia = a.forward(x)
    qb, db, ib = b.forward(x, row_chunk=100)
    assert np.array_equal(ia, ib)
    assert rel_err(b.embed_avg, a.embed_avg) <= 1e-6 and rel_err(b.cluster_size, a.cluster_size) <= 1e-6


def test_oracle_tie_breaks_to_lowest_index_and_eval_is_pure():
    g = load_golden("ties_eval")
    o = QuantizeOracle(64, 64, embed=g["embed0"])
    o.training = False
    before = o.state()
    _, _, ind = o.forward(g["x0"])
    assert np.all(ind[0:8] == 2) and np.all(ind[8:16] == 17)     # duplicates -> lowest index
    after = o.state()
    for k in before:
        assert np.array_equal(before[k], after[k])                # eval never touches buffers


def test_oracle_rejects_bad_inputs():
    o = QuantizeOracle(64, 32)
    with pytest.raises(TypeError):
        o.forward(np.zeros((4, 64), dtype=np.float64))
    with pytest.raises(ValueError):
        o.forward(np.zeros((4, 63), dtype=np.float32))
