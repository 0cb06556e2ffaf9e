#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules of the quantizer path next to the oracle so that they travel to the GPU box.

    python tools/fetch_ref.py            (also called by __graft_entry__.build())

`/root/reference` exists only in the build container; the GPU box gets a snapshot of this repository (git-ignored files
included).  This script copies the few Python files of the reference that define the path -- `vqvae.py` (Quantize,
VQVAE), `vqvae_deep.py` (the D = 256 fork) and the `distributed/` package they import -- byte for byte into
`oracle/_ref/` (listed in .gitignore: reference sources never enter this repository's history) and writes
`oracle/_ref/MANIFEST.json` with their SHA-256.  The committed `oracle/ref_manifest.json` pins the same digests, so the
GPU tests can prove that what they executed is the reference as shipped and not an edited copy.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's reference arms import from `oracle/_ref`; the product package
never does.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("VQB200_REFERENCE_DIR", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["vqvae.py", "vqvae_deep.py", "distributed/__init__.py", "distributed/distributed.py", "distributed/launch.py"]
PINNED = os.path.join(ROOT, "oracle", "ref_manifest.json")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def staged_ok():
    """True when oracle/_ref holds every file with the pinned digest."""
    if not os.path.exists(PINNED):
        return False
    pinned = json.load(open(PINNED))["files"]
    for rel, digest in pinned.items():
        p = os.path.join(DST, rel)
        if not os.path.exists(p) or sha256(p) != digest:
            return False
    return True


def fetch(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[fetch_ref] {SRC} not present (GPU box?): using the staged copy" if staged_ok()
                  else f"[fetch_ref] {SRC} not present and oracle/_ref is not staged")
        return staged_ok()
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if not os.path.exists(PINNED) or json.load(open(PINNED))["files"] != manifest:
        with open(PINNED, "w") as f:
            json.dump({"reference": "alehdaghi/vq-vae-2-pytorch", "files": manifest}, f, indent=1, sort_keys=True)
            f.write("\n")
    if verbose:
        print(f"[fetch_ref] staged {len(FILES)} reference files in oracle/_ref (digests pinned in oracle/ref_manifest.json)")
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
