"""CPU: the C-ABI library loads and exports every symbol include/b200render.h declares; struct layouts agree."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200render.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200r_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(rb):
    lib = C.CDLL(rb.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_mirror_covers_header(rb):
    from renderer_b200 import _abi
    assert set(declared_symbols()) == set(_abi.SYMBOLS)


def test_struct_sizes_match_c(rb):
    from renderer_b200 import _abi
    prog = r'''
#include <stdio.h>
#include "b200render.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %d\n", sizeof(b200r_vertex), sizeof(b200r_tri), sizeof(b200r_bvhnode),
  sizeof(b200r_light), sizeof(b200r_frame), sizeof(b200r_counters), sizeof(b200r_orbit), B200R_MAX_FRAMES_IN_FLIGHT); return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c"); open(c, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(t) for t in (_abi.Vertex, _abi.Tri, _abi.BvhNode, _abi.Light, _abi.Frame, _abi.Counters, _abi.Orbit)]
    assert sizes[:-1] == want and sizes[-1] == _abi.MAX_FRAMES_IN_FLIGHT
    assert sizes[0] == 28 and sizes[2] == 32      # = reference Vertex / CacheFriendlyBVHNode


def test_no_device_means_failure_not_fallback(rb):
    """On a box without a B200 the product must refuse to render (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rb.RendererError):
        rb.Renderer(0)


def test_product_never_touches_the_oracle():
    """The shipped package must not import/link anything under oracle/."""
    pkg = os.path.join(ROOT, "renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle_port" not in text and "liboracle" not in text and "from oracle" not in text \
                    and "import oracle" not in text, os.path.join(dirpath, fn)
