"""The C-ABI boundary without a GPU: the library builds for sm_100a, loads, exports every symbol
include/shx.h declares, mirrors the reference's parameter defaults, and refuses to run without a
device (there is no CPU fallback to fall into)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import simplehydrology_b200 as shx
from simplehydrology_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return shx.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "shx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shx_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/shx.h but not exported by libshx.so"


def test_no_undeclared_public_symbols(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", shx.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (shx_[a-z_0-9]+)", out))
    assert exported == set(declared_symbols())


def test_signatures_have_no_cuda_or_torch_types():
    text = open(os.path.join(ROOT, "include", "shx.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # prose may mention them, declarations may not
    for banned in ("cudaStream_t", "torch", "at::", "cuda_runtime"):
        assert banned not in code
    # compiles as plain C
    r = subprocess.run(["gcc", "-std=c11", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "shx.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_struct_layouts_match_the_reference(lib):
    assert C.sizeof(shx.Params) == 10 * 4 + 4 * 4
    assert shx.CELL_DTYPE.itemsize == 32  # quad::cell, cellpool.h:207-220
    assert shx.DROP_DTYPE.itemsize == 32  # Drop (28 B, water.h:12-39) + status word
    assert C.sizeof(shx.Stats) == 16 * 8


def test_default_params_are_the_reference_statics(lib):
    import numpy as np
    p = shx.default_params(4)
    f32 = np.float32
    assert (f32(p.evapRate), f32(p.depositionRate), f32(p.minVol), p.maxAge) == (f32(0.001), f32(0.1), f32(0.01), 500.0)  # water.h:43-46
    assert (p.entrainment, p.gravity, p.momentumTransfer) == (10.0, 1.0, 1.0)  # water.h:48-50
    assert (f32(p.lrate), f32(p.maxdiff), f32(p.settling)) == (f32(0.1), f32(0.01), f32(0.8))  # world.h:42-44
    assert (p.mapscale, p.tilesize, p.mapsize, p.lodsize) == (80, 512, 4, 1)  # cellpool.h:165-178


def test_bad_arguments_are_rejected_before_touching_cuda(lib):
    h = C.c_void_p()
    assert lib.shx_create(C.byref(h), None, None) == -1
    p = shx.default_params(1)
    p.lodsize = 2
    assert lib.shx_create(C.byref(h), C.byref(p), None) == -1 and b"lodsize" in lib.shx_last_error()
    p = shx.default_params(1)
    cfg = shx.Config()
    lib.shx_default_config(C.byref(cfg))
    cfg.row0, cfg.row1 = 100, 50
    assert lib.shx_create(C.byref(h), C.byref(p), C.byref(cfg)) == -1
    assert lib.shx_erode(None, 1, 0, None) == -1
    assert lib.shx_upload(None, None, 0) == -1


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(shx.ShxError) as e:
        shx.World(mapsize=1)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(shx, "_lib", None)
    monkeypatch.setattr(shx, "LIB_PATH", os.path.join(ROOT, "does_not_exist", "libshx.so"))
    with pytest.raises(shx.ShxError):
        shx.lib()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under simplehydrology_b200/ or include/ may reference it"""
    for base in ("simplehydrology_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "import orc" not in text and "shx_oracle" not in text and "libshx_ref" not in text, f


def test_host_adaptor_compiles_against_the_c_abi(lib):
    exe = B.build_host_example(force=True)
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


GL_TU = r"""
#include <GL/gl.h>
#define SHX_WITH_GL
#include "simplehydrology_b200/host/shx_world.hpp"
#include "simplehydrology_b200/host/shx_gl.hpp"
// one frame of SimpleHydrology.cpp:322-335 with the vertex pool's VBO and the tree instance buffer mapped into CUDA
int frame(shx::Bridge& b, unsigned vbo_name, unsigned instance_buffer_name) {
  shx::GLBuffer vbo(vbo_name), trees(instance_buffer_name);
  { auto m = vbo.map(); b.update_vertices_device(m.as<float>()); }
  { auto m = trees.map(); return (int)b.tree_models_device(m.as<float>(), m.bytes / 64); }
}
"""


def test_gl_interop_option_compiles(tmp_path):
    """SHX_WITH_GL (source/vertexpool.h:155-172: the vertex pool is one GL buffer): the registration / mapping wrapper
    compiles against the CUDA toolkit's cuda_gl_interop.h; the image has no GL headers, so the two typedefs that
    header needs come from tests/gl_stub"""
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_gl_interop.h")):
        pytest.skip("no CUDA toolkit headers")
    src = tmp_path / "gl_tu.cpp"
    src.write_text(GL_TU)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++17", "-Wall", "-c", str(src), "-I", ROOT, "-I", os.path.join(ROOT, "tests", "gl_stub"), "-I", cuda_inc,
                        "-o", str(tmp_path / "gl_tu.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run(["nm", "-C", str(tmp_path / "gl_tu.o")], capture_output=True, text=True).stdout
    assert "cudaGraphicsGLRegisterBuffer" in out and "shx_vertex_fill" in out and "shx_veg_tree_models" in out
