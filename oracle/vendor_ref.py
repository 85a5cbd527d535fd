"""TEST / BASELINE INFRASTRUCTURE ONLY - recipe that makes the UNMODIFIED reference available on the GPU box.

The reference is pure Python, so there is nothing to compile; what `oracle/_ref` holds for a compiled reference (a built
.so) is here ONE archive (`oracle/_ref/reference_py.tar.gz`) of the reference's own `frido/` and `taming/` packages and the
shipped `configs/frido/` YAMLs, packed where they lie under /root/reference and unpacked into a temp directory at run
time.  `oracle/_ref/` is git-ignored (never part of the
history: no reference source is committed) but travels to the GPU box with the snapshot like a built .so, so that

  * `bench.py --impl reference` times the reference's OWN modules on the box's host cores (`cpu_baseline.kind` =
    "reference"), and
  * `bench.py` can time the reference's own eager PyTorch path on the B200 as the software baseline of SURVEY.md §8(d).

Nothing under frido_b200/ imports it.  Run:  python oracle/vendor_ref.py   (also called by __graft_entry__.build()).
"""
import os
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FRIDO_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DST, "reference_py.tar.gz")
_SKIP = ("__pycache__", ".pyc", ".ckpt", ".pth", ".pt", ".png", ".jpg", ".ttf")


def vendor(verbose=True):
    """Pack <reference>/{frido,taming,configs/frido} into oracle/_ref/reference_py.tar.gz (one build artefact, like a .so)."""
    if not os.path.isdir(os.path.join(SRC, "frido")):
        if verbose:
            print(f"vendor_ref: {SRC} not present - keeping {ARCHIVE if os.path.exists(ARCHIVE) else 'nothing'}")
        return os.path.exists(ARCHIVE)
    os.makedirs(DST, exist_ok=True)
    n = 0
    with tarfile.open(ARCHIVE, "w:gz") as tar:
        for sub in ("frido", "taming", os.path.join("configs", "frido")):
            for dirpath, dirnames, files in os.walk(os.path.join(SRC, sub)):
                dirnames[:] = sorted(d for d in dirnames if d != "__pycache__")
                for f in sorted(files):
                    if f.endswith(_SKIP):
                        continue
                    full = os.path.join(dirpath, f)
                    tar.add(full, arcname=os.path.relpath(full, SRC))
                    n += 1
    if verbose:
        print(f"vendor_ref: {n} files -> {ARCHIVE} ({os.path.getsize(ARCHIVE) >> 10} KiB)")
    return True


def unpack():
    """Extract the archive into a per-user temp directory (once per archive state); returns that directory or None."""
    if not os.path.exists(ARCHIVE):
        return None
    st = os.stat(ARCHIVE)
    out = os.path.join(tempfile.gettempdir(), f"frido_ref_{os.getuid()}_{int(st.st_mtime)}_{st.st_size}")
    if not os.path.isdir(os.path.join(out, "frido")):
        tmp = out + f".part{os.getpid()}"
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            tar.extractall(tmp, filter="data")
        try:
            os.rename(tmp, out)
        except OSError:  # another process won the race
            pass
    return out


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
