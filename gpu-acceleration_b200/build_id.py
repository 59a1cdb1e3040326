"""Source hash of everything libb200msm.so is compiled from (csrc/*.cu, *.cuh, *.inc + include/*.h).

The Makefile embeds it in the library (`b200msm_build_id()`); `b200msm.load_library()` recomputes it from the
checkout and refuses (or rebuilds) a library that was linked from different sources, so the binary that is tested
is always the tree that is committed."""
import glob
import hashlib
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def source_files():
    pats = ["csrc/*.cu", "csrc/*.cuh", "csrc/*.inc", "../include/*.h"]
    out = []
    for p in pats:
        out += glob.glob(os.path.join(HERE, p))
    return sorted(out, key=lambda f: os.path.basename(f))


def build_id() -> str:
    h = hashlib.sha256()
    for f in source_files():
        h.update(os.path.basename(f).encode())
        h.update(b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
        h.update(b"\0")
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(build_id())
