"""Make the UNMODIFIED reference importable on the GPU box: copies the Python sources the decoder path needs from
/root/reference into baseline/_ref/ (git-ignored, not gpurun-ignored, so it travels with the snapshot; never committed).

The reference is not a pip package (no setup.py / pyproject at its root), so `pip install --target baseline/_ref
/root/reference` has nothing to build; this copy is the equivalent.  bench.py --impl reference imports it with the
import shims of SURVEY.md Appendix C.  Run in the build container only (called by __graft_entry__.build())."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
WANT = ("models", "utils", "datasets")


def install(force=False):
    if not os.path.isdir(os.path.join(SRC, "models")):
        return False
    if os.path.isdir(os.path.join(DST, "models")) and not force:
        return True
    for d in WANT:
        dst = os.path.join(DST, d)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, d), dst, ignore=lambda _d, names: [n for n in names if not (n.endswith(".py") or "." not in n)])
    with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
        f.write("verbatim copy of the .py files of /root/reference/{models,utils,datasets} (V-DETR/V-DETR); not part of this repository\n")
    return True


if __name__ == "__main__":
    print("installed" if install("--force" in sys.argv) else "no /root/reference here")
