"""Place the UNMODIFIED reference under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).

    python scripts/install_reference.py [--source /root/reference]

The reference ships no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to
build ("neither 'setup.py' nor 'pyproject.toml' found"); it is a plain source tree that `benchmark.py` runs from its own
root.  This script therefore copies exactly the parts the training path imports or reads -- deepxde/, src/, trainer.py,
benchmark.py and the solution tables ref/*.dat -- byte for byte, and writes a manifest with their sha256 so a test can show
the copy is unmodified.  Nothing under baseline/_ref is tracked by git and nothing in pinnacle_b200/ imports it: it is
the comparator (tests/refshim.py, bench.py --impl reference), never the product.
"""
import argparse
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ["deepxde", "src", "ref", "trainer.py", "benchmark.py", "LICENSE", "requirements.txt"]


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="/root/reference")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.source, "deepxde")):
        raise SystemExit(f"no reference tree at {a.source}")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    manifest = {}
    for part in PARTS:
        s, d = os.path.join(a.source, part), os.path.join(DEST, part)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.exists(s):
            shutil.copy2(s, d)
    for dp, dn, fn in os.walk(DEST):
        for f in fn:
            p = os.path.join(dp, f)
            os.chmod(p, 0o644)
            manifest[os.path.relpath(p, DEST)] = sha(p)
        os.chmod(dp, 0o755)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": a.source, "files": manifest}, f, indent=0, sort_keys=True)
    print(f"installed {len(manifest)} files under {DEST}")


if __name__ == "__main__":
    main()
