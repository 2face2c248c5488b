"""Recipe for oracle/_ref — the REAL reference package, importable — TEST INFRASTRUCTURE ONLY.

The reference (`/root/reference`, Ahnsun/merlin) is pure Python; "building" it means making it importable:

  1. copy `/root/reference/mmgpt` to `oracle/_ref/mmgpt` (git-ignored, NOT gpurun-ignored: it travels to the
     GPU box with the snapshot exactly like the built `.so`; the reference's sources never enter history);
  2. restore the one missing `)` of `mmgpt/utils/constants.py:72` (the tree as scraped opens
     `CONVERSATION_DATA = dict(` at :25 and never closes it — SURVEY.md §0 fact 3).  Nothing else is touched.

`oracle/refmodel.py` imports the copy; `tests/`, `bench.py --impl reference` / `cpu_baseline` and
`oracle/make_golden.py` are its only users.  Run by `__graft_entry__.build()` whenever /root/reference is
mounted (the build container); on the GPU box the prebuilt copy is used as it arrived.

Usage: python oracle/build_ref.py
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
OUT = os.path.join(HERE, "_ref")


def tree_sha(root: str) -> str:
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    src = os.path.join(REFERENCE, "mmgpt")
    if not os.path.isdir(src):
        if os.path.isdir(os.path.join(OUT, "mmgpt")):
            return OUT                      # GPU box: use the copy that travelled with the snapshot
        raise FileNotFoundError(f"{src} is not mounted and {OUT} has not been built")
    stamp = os.path.join(OUT, "PROVENANCE.json")
    want = tree_sha(src)
    if not force and os.path.exists(stamp) and json.load(open(stamp)).get("source_sha256") == want:
        return OUT
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    shutil.copytree(src, os.path.join(OUT, "mmgpt"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.jpg", "*.png"))
    for d, _, files in os.walk(OUT):
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)
    p = os.path.join(OUT, "mmgpt", "utils", "constants.py")
    lines = open(p).read().split("\n")
    lines.insert(71, ")")                   # closes CONVERSATION_DATA = dict(  (constants.py:25)
    open(p, "w").write("\n".join(lines))
    compile(open(p).read(), p, "exec")      # the patched file must parse
    commit = None
    sub = os.path.join(REFERENCE, ".SUBMODULES.json")
    if os.path.exists(sub):
        try:
            commit = json.load(open(sub))
        except Exception:
            commit = None
    json.dump({"source": src, "source_sha256": want, "patch": "inserted ')' as line 72 of mmgpt/utils/constants.py",
               "submodules": commit}, open(stamp, "w"), indent=1)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
