#!/usr/bin/env python
"""Build libanorag_b200.so in-tree with nvcc for sm_100a (the only target).

    python ano-rag_b200/csrc/build.py [--force] [--verbose]

The shared object is git-ignored but travels to the GPU box with the gpurun
snapshot.  Objects are rebuilt only when a source or header is newer.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["common.cu", "dense.cu", "bm25.cu", "bm25_batch.cu", "fuse.cu", "tfidf.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "bm25.cuh", os.path.join(ROOT, "include", "anorag_b200.h")]
LIB = os.path.join(HERE, "libanorag_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-Xptxas", "-v", "--fmad=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libanorag_b200 cannot be built")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    hdrs = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    objs = []
    logs = []
    jobs = []
    for src in SOURCES:
        sp = os.path.join(HERE, src)
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [sp, __file__] + hdrs):
            jobs.append((src, [nvcc] + ARCH + FLAGS + ["-c", sp, "-o", obj]))

    def compile_one(job):
        src, cmd = job
        return src, cmd, subprocess.run(cmd, capture_output=True, text=True)

    if jobs:                                              # the translation units are independent: compile them in parallel
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            for src, cmd, r in pool.map(compile_one, jobs):
                logs.append(r.stderr)
                if verbose:
                    print(" ".join(cmd)); print(r.stderr)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if force or _newer(LIB, objs):
        # the shared CUDA runtime: torch has already mapped libcudart.so.12 when the library is loaded (_native.py imports
        # torch first), and the shipped object then carries no copy of the runtime's entry-point tables
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "shared"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if logs:
        with open(os.path.join(HERE, "ptxas_info.log"), "w") as f:
            f.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
