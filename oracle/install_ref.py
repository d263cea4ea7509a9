"""TEST / BENCH INFRASTRUCTURE ONLY — makes the UNMODIFIED reference available to `bench.py --impl reference`.

The driver's contract asks for `pip install --target baseline/_ref /root/reference`.  That cannot work: the reference
ships no setup.py / pyproject.toml (it is a script tree: main.py + src/), so pip has nothing to build.  Recorded in
DESIGN.md.  The recipe here is therefore a plain copy of the reference's Python tree (main.py + src/, ~400 KB, no
data blobs) into `baseline/_ref/` — git-ignored (never enters the history), NOT gpurun-ignored (it travels to the GPU
box with the snapshot, like the built .so).  The six third-party modules the reference imports but this image lacks
are stubbed at import time by oracle/ref_shim.py; not one reference file is edited.

    python oracle/install_ref.py            # /root/reference -> baseline/_ref ; prints the content hash
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("FEDCOLA_REFERENCE_SOURCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def tree_hash(root):
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()[:16]


def installed():
    return os.path.isdir(os.path.join(DST, "src", "server"))


def install(force=False):
    """Copy main.py + src/ (Python files only) if the source tree exists. Returns the install dir or None."""
    if not os.path.isdir(os.path.join(SRC, "src")):
        return DST if installed() else None
    if installed() and not force and tree_hash(os.path.join(DST, "src")) == tree_hash(os.path.join(SRC, "src")):
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, "src"), os.path.join(DST, "src"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.sh"))
    for f in ("main.py", "LICENSE", "requirments.txt"):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"unmodified copy of {SRC} (main.py + src/), sha256[:16] of src/**/*.py = {tree_hash(os.path.join(DST, 'src'))}\n"
                "made by oracle/install_ref.py; git-ignored; used only by bench.py's reference arms\n")
    return DST


if __name__ == "__main__":
    d = install(force="--force" in sys.argv)
    print(d, tree_hash(os.path.join(d, "src")) if d else "reference source not found")
