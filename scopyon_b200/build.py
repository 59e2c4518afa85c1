"""Build the CUDA shared library in-tree (``scopyon_b200/libscopyon_b200.so``).

``nvcc`` cross-compiles for sm_100a without a GPU, so this runs in the build
container; the resulting ``.so`` is git-ignored and travels to the GPU box with the
working tree.  ``python -m scopyon_b200.build`` or ``__graft_entry__.build()``.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscopyon_b200.so")
OBJ_DIR = os.path.join(HERE, "_build")
SOURCES = ["psf.cu", "render.cu", "particles.cu", "detector.cu", "gaussian_tc.cu", "frames.cu", "spots.cu",
           "host_frames.cpp"]       # .cpp: host-only code (nvcc hands it to g++)
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: scopyon_b200 has no CPU fallback and cannot be built without CUDA")
    return exe


def _digest(paths):
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _deps(src):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "scopyon_b200.h"))
    return [src] + headers


def build_library(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` for sm_100a and link the shared library."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    sources = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")
        stamp = obj + ".sha"
        digest = _digest(_deps(src))
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
            continue
        jobs.append((src, obj, stamp, digest))

    def compile_one(job):
        src, obj, stamp, digest = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for {}:\n{}\n{}".format(src, res.stdout, res.stderr))
        with open(stamp, "w") as f:
            f.write(digest)
        return res.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for log in pool.map(compile_one, jobs):
            if verbose and log:
                print(log, file=sys.stderr)

    objs = [os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(s))[0] + ".o") for s in sources]
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcuda"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n{}\n{}".format(res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
