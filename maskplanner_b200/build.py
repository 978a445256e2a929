"""Builds libmaskplanner_b200.so in-tree with nvcc for sm_100a (and nothing else).

    python -m maskplanner_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/maskplanner_b200.h); it is loaded with ctypes
by maskplanner_b200._cabi.  The .so is git-ignored but travels to the GPU box with the repo
snapshot.  Object files are cached per source under maskplanner_b200/_build/.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libmaskplanner_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "-Xptxas", "-v", "-Xcudafe", "--diag_suppress=177"]
if os.environ.get("MPB_MBAR_DEBUG", "0") == "1":      # developer build: timed-out mbarrier waits record their location (tc_common.cuh)
    NVCC_FLAGS.append("-DMPB_MBAR_DEBUG")


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmaskplanner_b200.so cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def source_hash():
    """sha256 over every csrc/ source, header and the public C header: identifies the build the .so came from."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(HERE, "..", "include", "maskplanner_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH + NVCC_FLAGS).encode())
    return h.hexdigest()


HASH_FILE = LIB + ".srchash"


def is_stale():
    """True when the .so is missing or was built from different sources (mtime-independent: snapshots reset mtimes)."""
    if not os.path.exists(LIB) or not os.path.exists(HASH_FILE):
        return True
    with open(HASH_FILE) as f:
        return f.read().strip() != source_hash()


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "maskplanner_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = _deps_mtime()
    force = force or (os.path.exists(HASH_FILE) and is_stale())   # a stale hash means mtimes cannot be trusted either
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t)
        jobs.append((src, obj, stale))

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return ""
        cmd = [nvcc] + ARCH + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(obj + ".ptxas.txt", "w") as f:
            f.write(r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(jobs) or 1)) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    objs = [j[1] for j in jobs]
    if force or any(j[2] for j in jobs) or not os.path.exists(LIB):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(HASH_FILE, "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
