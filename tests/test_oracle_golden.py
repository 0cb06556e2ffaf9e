"""Pin the CPU oracle (oracle/quantize_oracle.py) against outputs of the reference module.

The fixtures were produced by tests/golden/make_golden.py, which executes the reference
`Quantize` (/root/reference/vqvae.py:28-78) on seeded inputs.  CPU-only.
"""
import numpy as np
import pytest

from helpers import REL_TOL, col_rel_err, golden_names, load_golden, rel_err
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
        q, diff, ind = o.forward(x)
        assert ind.dtype == np.int64 and ind.shape == x.shape[:-1]
        ndiff, nbad, _ = tie_tolerant_index_mismatches(x, embed_before, ind, g[f"ind{s}"])
        assert nbad == 0, f"{name} step {s}: {nbad} index mismatches beyond fp32 near-ties"
        if ndiff == 0:
            assert rel_err(q, g[f"quantize{s}"]) <= REL_TOL
            assert abs(float(diff) - float(g[f"diff{s}"])) <= REL_TOL * abs(float(g[f"diff{s}"])) + 1e-12
            assert rel_err(o.cluster_size, g[f"cluster_size{s + 1}"]) <= REL_TOL
            assert col_rel_err(o.embed_avg, g[f"embed_avg{s + 1}"]) <= REL_TOL
            assert col_rel_err(o.embed, g[f"embed{s + 1}"]) <= 4 * REL_TOL
        # keep both trajectories on the reference's state so later steps stay comparable
        o.load(g[f"embed{s + 1}"], g[f"cluster_size{s + 1}"], g[f"embed_avg{s + 1}"])
        if f"xgrad{s}" in g:
            codes = embed_before.T[g[f"ind{s}"]]
            gx = o.backward(x, codes, g[f"gq{s}"], float(g[f"gd{s}"]))
            assert rel_err(gx, g[f"xgrad{s}"]) <= REL_TOL


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
