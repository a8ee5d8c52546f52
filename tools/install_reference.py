#!/usr/bin/env python
"""Put the UNMODIFIED reference (colehurwitz/llm_bci) where bench.py's reference arm can run it on the GPU box.

The reference is a plain Python tree (no setup.py / pyproject.toml: `pip install /root/reference` has nothing to build),
so "installing" it means copying the files its NDT1 path imports -- models/, utils/, data_utils/, configs/, vocab.json --
byte for byte from /root/reference into baseline/_ref/.  That directory is git-ignored (reference sources never enter this
repository's history) but NOT gpurun-ignored, so it travels to the GPU box with the snapshot, where /root/reference does
not exist.  baseline/reference_manifest.json (committed) holds the sha256 of every copied file; bench.py verifies the copy
against it before timing, which is what "unmodified" means in its `cpu_baseline.kind = "reference"`.

Run in the build container:   python tools/install_reference.py        (also done by __graft_entry__.build())
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
MANIFEST = os.path.join(ROOT, "baseline", "reference_manifest.json")
PARTS = ["models", "utils", "data_utils", "configs", "vocab.json", "LICENSE"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def tree_hashes(base):
    out = {}
    for part in PARTS:
        p = os.path.join(base, part)
        if os.path.isfile(p):
            out[part] = sha256(p)
        for d, _, files in os.walk(p):
            if "__pycache__" in d:
                continue
            for f in sorted(files):
                if f.endswith(".pyc"):
                    continue
                full = os.path.join(d, f)
                out[os.path.relpath(full, base)] = sha256(full)
    return dict(sorted(out.items()))


def verify(base=DST):
    """True if `base` holds exactly the files of the manifest with the same contents."""
    if not (os.path.exists(MANIFEST) and os.path.isdir(base)):
        return False
    want = json.load(open(MANIFEST))["files"]
    return tree_hashes(base) == want


def install(write_manifest=True):
    if not os.path.isdir(SRC):
        return verify()
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    if write_manifest:
        files = tree_hashes(SRC)
        json.dump({"source": "colehurwitz/llm_bci as mounted at /root/reference", "files": files}, open(MANIFEST, "w"), indent=1)
    return verify()


if __name__ == "__main__":
    ok = install()
    print("baseline/_ref", "installed and verified" if ok else "NOT available", f"({len(tree_hashes(DST)) if os.path.isdir(DST) else 0} files)")
    sys.exit(0 if ok else 1)
