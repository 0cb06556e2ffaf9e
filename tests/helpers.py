"""Shared helpers for the test-suite (fixture loading, tolerant comparisons)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-5   # BASELINE.json: quantize / diff / EMA buffers within 1e-5 relative in fp32


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """max |a-b| relative to the scale of b (max-norm): robust for buffers that mix 1e5 and 1e-2."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-30)
    return float(np.max(np.abs(a - b)) / denom) if b.size else 0.0


def col_rel_err(a, b):
    """per-codebook-column relative error for [D,K] buffers (each code judged on its own scale)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    num = np.max(np.abs(a - b), axis=0)
    den = np.maximum(np.max(np.abs(b), axis=0), 1e-30)
    return float(np.max(num / den))
