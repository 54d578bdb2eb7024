#!/usr/bin/env python3
"""Build the UNMODIFIED reference renderer headless -> oracle/_ref/  (TEST INFRASTRUCTURE ONLY).

The reference (ttsiodras/renderer, mounted read-only at /root/reference) hard-codes its
resolution and ray-tracer features as #defines, so every (W, H, feature) combination is a
separate binary.  This recipe

  1. copies /root/reference/src to a scratch dir under /tmp (never into the repo),
  2. applies the per-config patch list of SURVEY.md §8c with exact-match substitutions
     (WIDTH/HEIGHT, REFLECTIONS, AMBIENT_OCCLUSION/AMBIENT_SAMPLES [+ deterministic rand shim],
     MLAA),
  3. compiles it with g++ against the headless stub in oracle/sdl_stub/ and the vendored
     lib3ds C files compiled where they lie,
  4. writes ONLY the binary to oracle/_ref/bin/renderer_<tag>.

Two flavours:
  strict (default): -O3 -fopenmp -DNDEBUG, no fast-math, no -march  -> the PARITY oracle
  --fast          : the reference's own configure.ac flags (+SIMD_SSE/SSE2 as its configure
                    would define on x86-64)                           -> the TIMING baseline

Nothing under oracle/ is used by the product; see oracle/README.md.
"""
import argparse
import concurrent.futures as cf
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RENDERER_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
STUB = os.path.join(HERE, "sdl_stub")

CXX_SOURCES = ["renderer", "Base3d", "BVH", "Camera", "Keyboard", "Light", "Loader",
               "Rasterizers", "Raytracer", "Screen", "Wu"]

STRICT_FLAGS = ["-O3", "-fopenmp", "-DNDEBUG", "-w"]
FAST_FLAGS = ["-O3", "-ffast-math", "-funsafe-math-optimizations", "-mtune=native", "-msse",
              "-msse2", "-mssse3", "-mrecip", "-mfpmath=sse", "-fomit-frame-pointer",
              "-fopenmp", "-DNDEBUG", "-w", "-DSIMD_SSE=1", "-DSIMD_SSE2=1"]


def tag_of(a):
    t = f"{a.w}x{a.h}"
    if a.no_reflections:
        t += "_norefl"
    if a.ao:
        t += f"_ao{a.ao}"
    if a.mlaa:
        t += "_mlaa"
    if a.fast:
        t += "_fast"
    if getattr(a, "libc_rand", False):
        t += "_libcrand"
    return t


def sub_exact(path, pattern, repl, count=1):
    """Regex-substitute and insist the pattern matched exactly `count` times."""
    with open(path, "r", encoding="latin-1") as f:
        s = f.read()
    s2, n = re.subn(pattern, repl, s, flags=re.M)
    if n != count:
        raise SystemExit(f"patch failed on {path}: {pattern!r} matched {n} times, wanted {count}")
    with open(path, "w", encoding="latin-1") as f:
        f.write(s2)


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise SystemExit(1)


def build_lib3ds(objdir):
    """The 18 vendored lib3ds C files, compiled from where they lie (needed only to link)."""
    lib = os.path.join(OUT, "obj", "lib3ds.a")
    if os.path.exists(lib):
        return lib
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(REF, "lib3ds-1.3.0", "lib3ds", "*.c")))
    objs = []
    with cf.ThreadPoolExecutor(8) as ex:
        futs = []
        for s in srcs:
            o = os.path.join(objdir, "l3_" + os.path.basename(s)[:-2] + ".o")
            objs.append(o)
            futs.append(ex.submit(run, ["gcc", "-O2", "-w", "-I", os.path.join(REF, "lib3ds-1.3.0"),
                                        "-c", s, "-o", o]))
        for f in futs:
            f.result()
    run(["ar", "rcs", lib] + objs)
    return lib


def build(a):
    tag = tag_of(a)
    bindir = os.path.join(OUT, "bin")
    os.makedirs(bindir, exist_ok=True)
    exe = os.path.join(bindir, "renderer_" + tag)
    if os.path.exists(exe) and not a.force and \
            os.path.getmtime(exe) > max(os.path.getmtime(__file__),
                                        os.path.getmtime(os.path.join(STUB, "sdl_stub.cc")),
                                        os.path.getmtime(os.path.join(STUB, "SDL.h"))):
        print(exe)
        return exe
    if not os.path.isdir(os.path.join(REF, "src")):
        raise SystemExit(f"reference sources not found under {REF}; cannot build {exe}")

    work = tempfile.mkdtemp(prefix="oracle_ref_")
    try:
        src = os.path.join(work, "src")
        shutil.copytree(os.path.join(REF, "src"), src)
        for root, _, files in os.walk(src):
            for fn in files:
                os.chmod(os.path.join(root, fn), 0o644)
        with open(os.path.join(work, "config.h"), "w") as f:
            f.write("#define HAVE_GETOPT_H 1\n#define USE_OPENMP 1\n")
            if a.mlaa:
                f.write("#define MLAA_ENABLED 1\n")

        # --- the per-config patch list ---
        sub_exact(os.path.join(src, "Defines.h"), r"^#define WIDTH\s+800\s*$", f"#define WIDTH {a.w}")
        sub_exact(os.path.join(src, "Defines.h"), r"^#define HEIGHT\s+600\s*$", f"#define HEIGHT {a.h}")
        rt = os.path.join(src, "Raytracer.cc")
        if a.no_reflections:
            sub_exact(rt, r"^#define REFLECTIONS\s*$", "//#define REFLECTIONS")
        if a.ao:
            sub_exact(rt, r"^//#define AMBIENT_OCCLUSION\s*$", "#define AMBIENT_OCCLUSION")
            sub_exact(rt, r"^#define AMBIENT_SAMPLES\s+32\s*$", f"#define AMBIENT_SAMPLES {a.ao}")
            if not a.libc_rand:
                # the single permitted semantic delta: seed the per-pixel random stream
                sub_exact(rt, r"^(\s*for\(int x=xStarting; x<iOnePastEndingX; x\+\+\) \{)\s*$",
                          r"\1 oracle_seed(x, y);")
                # ... and draw from it instead of the process-global, racy libc rand() (3 call sites, :393-395)
                sub_exact(rt, r"float\(rand\(\)-RAND_MAX/2\)", "float(oracle_rand()-RAND_MAX/2)", count=3)
            # (--libc-rand: AO exactly as the reference has it, libc rand() and all - for the statistical comparison only)

        flags = FAST_FLAGS if a.fast else STRICT_FLAGS
        inc = ["-I", work, "-I", STUB, "-I", os.path.join(REF, "lib3ds-1.3.0")]
        units = list(CXX_SOURCES) + (["MLAA"] if a.mlaa else [])
        objs = []
        with cf.ThreadPoolExecutor(8) as ex:
            futs = []
            for u in units:
                o = os.path.join(work, u + ".o")
                objs.append(o)
                extra = []
                futs.append(ex.submit(run, ["g++", "-std=gnu++14"] + flags + extra + inc +
                                      ["-c", os.path.join(src, u + ".cc"), "-o", o]))
            o = os.path.join(work, "sdl_stub.o")
            objs.append(o)
            futs.append(ex.submit(run, ["g++", "-O2", "-w", "-I", STUB, "-c",
                                        os.path.join(STUB, "sdl_stub.cc"), "-o", o]))
            lib = build_lib3ds(work)
            for f in futs:
                f.result()
        run(["g++", "-fopenmp", "-o", exe] + objs + [lib, "-lm"])
    finally:
        shutil.rmtree(work, ignore_errors=True)
    print(exe)
    return exe


def stage_models():
    """Copy the reference's model files to a writable, git-ignored dir that travels to the GPU
    box (the reference writes its .bvh cache next to the model, Raytracer.cc:747)."""
    dst = os.path.join(OUT, "models")
    os.makedirs(dst, exist_ok=True)
    srcdir = os.path.join(REF, "3D-Objects")
    if not os.path.isdir(srcdir):
        return dst
    for p in sorted(glob.glob(os.path.join(srcdir, "*"))):
        if p.endswith((".ply", ".tri")):
            q = os.path.join(dst, os.path.basename(p))
            if not os.path.exists(q) or os.path.getsize(q) != os.path.getsize(p):
                shutil.copyfile(p, q)
                os.chmod(q, 0o644)
    return dst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=800)
    ap.add_argument("--h", type=int, default=600)
    ap.add_argument("--no-reflections", action="store_true")
    ap.add_argument("--ao", type=int, default=0, help="enable AMBIENT_OCCLUSION with N samples")
    ap.add_argument("--mlaa", action="store_true")
    ap.add_argument("--fast", action="store_true", help="reference's own flags (timing baseline)")
    ap.add_argument("--libc-rand", action="store_true",
                    help="with --ao: keep the reference's libc rand() (not reproducible in parallel; statistics only)")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--stage-models", action="store_true")
    a = ap.parse_args()
    if a.stage_models:
        print(stage_models())
        return
    build(a)


if __name__ == "__main__":
    main()
