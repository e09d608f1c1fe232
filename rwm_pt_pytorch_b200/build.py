"""Build librwmpt.so (hand-written sm_100a CUDA + the C ABI of include/rwmpt.h) in-tree with nvcc.

    python -m rwm_pt_pytorch_b200.build [--force]

No torch headers are involved (the ABI is plain C), so the library is independent of the torch wheel's
CUDA version.  nvcc cross-compiles without a GPU; the target-family translation units build in parallel.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.environ.get("RWMPT_CSRC", os.path.join(PKG, "csrc"))  # RWMPT_CSRC: build an alternative source tree (A/B tests)
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "librwmpt.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-I", INCLUDE,
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: librwmpt.so cannot be built")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), tag: str = "", only: tuple = ()) -> str:
    """Compile every .cu under csrc/ for sm_100a and link librwmpt.so; returns the library path.
    `defines` / `tag` build an experimental variant (librwmpt_<tag>.so) with extra -D flags."""
    global OBJ, LIB
    if tag:
        OBJ = os.path.join(PKG, "build", tag)
        LIB = os.path.join(PKG, f"librwmpt_{tag}.so")
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")) + [__file__]
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    nvcc = _nvcc()
    jobs = []
    objs = []
    main_obj = os.path.join(PKG, "build")
    for src in sources:
        stem = os.path.basename(src)[:-3]
        if tag and only and not any(o in stem for o in only):
            objs.append(os.path.join(main_obj, stem + ".o"))  # variant build: untouched families come from the main build
            continue
        obj = os.path.join(OBJ, stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{p.stdout}\n{p.stderr}")
        return p.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for log in ex.map(compile_one, jobs):
                if verbose and log:
                    print(log)
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    NVCC_FLAGS.extend(a[7:] for a in sys.argv if a.startswith("--nvcc="))   # extra raw nvcc flags for A/B builds
    defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    tags = [a[6:] for a in sys.argv if a.startswith("--tag=")]
    only = tuple(x for a in sys.argv if a.startswith("--only=") for x in a[7:].split(","))  # with --tag: rebuild these TUs only
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, tag=tags[0] if tags else "", only=only)
    print(path)
