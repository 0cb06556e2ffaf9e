"""Loader of the UNMODIFIED reference modules staged in oracle/_ref  --  TEST INFRASTRUCTURE ONLY.

`tools/fetch_ref.py` (run by `__graft_entry__.build()` in the build container, where /root/reference is mounted)
copies the reference's `vqvae.py`, `vqvae_deep.py` and `distributed/` byte for byte into `oracle/_ref/` (git-ignored,
but it travels to the GPU box with the snapshot).  This module imports them from there -- the reference's own
`Quantize` / `VQVAE` classes, executed as they are (on CUDA in the `-m gpu` tests, on the host cores in bench.py's
reference arm) -- after checking every file against the SHA-256 digests pinned in the committed
`oracle/ref_manifest.json`.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's reference arms may import this file; the product package
(`vq_vae_2_pytorch_b200`) never does.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PINNED = os.path.join(HERE, "ref_manifest.json")

_cache = {}


class ReferenceUnavailable(RuntimeError):
    pass


def _sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def verify():
    """Raise ReferenceUnavailable unless oracle/_ref holds every pinned file, bit-identical to the reference."""
    if not os.path.exists(PINNED):
        raise ReferenceUnavailable("oracle/ref_manifest.json missing")
    pinned = json.load(open(PINNED))["files"]
    for rel, digest in pinned.items():
        p = os.path.join(REF_DIR, rel)
        if not os.path.exists(p):
            raise ReferenceUnavailable(f"oracle/_ref/{rel} is not staged: run `python tools/fetch_ref.py` (or "
                                       "__graft_entry__.build()) where /root/reference is mounted")
        if _sha256(p) != digest:
            raise ReferenceUnavailable(f"oracle/_ref/{rel} differs from the pinned reference digest")
    return sorted(pinned)


def available() -> bool:
    try:
        verify()
        return True
    except ReferenceUnavailable:
        return False


def load(name="vqvae"):
    """Import the staged reference module `vqvae` or `vqvae_deep` (its `import distributed` resolves to the staged package)."""
    if name in _cache:
        return _cache[name]
    verify()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        mod = importlib.import_module(name)
    finally:
        sys.dont_write_bytecode = old
    if os.path.dirname(os.path.abspath(mod.__file__)) != REF_DIR:
        raise ReferenceUnavailable(f"module {name} resolved to {mod.__file__}, not to the staged reference")
    _cache[name] = mod
    return mod
