"""Build libb200render.so (host plumbing + sm_100a kernels) and the b200renderer CLI, in-tree.

    python -m renderer_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU. Flags that matter for parity (DESIGN.md "parity"):
  device: -fmad=false (no FMA contraction), default -prec-div=true -prec-sqrt=true -ftz=false
  host  : -ffp-contract=off, no -ffast-math, no -march
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libb200render.so")
CLI = os.path.join(PKG, "b200renderer")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off",
                     "-Xptxas", "-v", "--expt-relaxed-constexpr"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-unused-function"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build failed: " + cmd[-1])
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "**", "*.h"), recursive=True) + \
        glob.glob(os.path.join(CSRC, "**", "*.cuh"), recursive=True) + \
        [os.path.join(ROOT, "include", "b200render.h"), os.path.abspath(__file__)]
    cus = sorted(glob.glob(os.path.join(CSRC, "cuda", "*.cu")))
    cpps = sorted(p for p in glob.glob(os.path.join(CSRC, "host", "*.cpp")) if not p.endswith("main.cpp"))
    jobs, objs = [], []
    for src in cus + cpps:
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(o)
        if force or _newer(o, [src] + headers):
            if src.endswith(".cu"):
                cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", o]
            else:
                cmd = ["g++"] + CXX_FLAGS + ["-c", src, "-o", o]
            jobs.append((cmd, o + ".log"))
    if jobs:
        with cf.ThreadPoolExecutor(min(8, len(jobs))) as ex:
            for out in ex.map(lambda j: _run(*j), jobs):
                if verbose:
                    print(out)
    if force or jobs or _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread"], os.path.join(OBJ, "link.log"))
    main = os.path.join(CSRC, "host", "main.cpp")
    if os.path.exists(main) and (force or _newer(CLI, [main, LIB] + headers)):
        _run(["g++"] + CXX_FLAGS + ["-pthread", "-o", CLI, main, "-L" + PKG, "-lb200render", "-Wl,-rpath,$ORIGIN"],
             os.path.join(OBJ, "cli.log"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
