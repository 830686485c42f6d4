"""The C-ABI shared library loads and exports every symbol include/celeste_cuda.h declares; without
a GPU every compute entry point fails LOUDLY (no CPU fallback).  CPU only: no compute calls."""
import ctypes as C
import os
import re

import pytest

import celeste_jl_b200 as cj
from celeste_jl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "celeste_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(celeste_[a-z0-9_]+)\s*\(", src)))


def test_library_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build()"


def test_exports_every_declared_symbol():
    lib = _lib.load()
    decl = _declared()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/celeste_cuda.h but not exported"
    assert sorted(_lib.EXPORTS) == decl


def test_struct_layouts_match_header():
    # sizes the C compiler produces for the header structs (x86-64 SysV)
    assert C.sizeof(_lib.celeste_image) == 48
    assert C.sizeof(_lib.celeste_patch) == 128


def test_errmsg_and_version():
    assert _lib.load().celeste_version() >= 100
    assert _lib.errmsg(0) == "ok"
    assert "no CPU fallback" in _lib.errmsg(_lib.CELESTE_ERR_NO_DEVICE)
    assert len(_lib.errmsg(12345)) < 61


def test_sass_is_sm100a_fp64():
    """The shipped cubin targets sm_100a and the hot kernel is FP64 (DFMA) code."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    funcs = [f for f in sass.split("Function : ") if f.startswith("_ZN7celeste12pixel_kernelILi2ELi2ELb0E")]
    assert len(funcs) == 1 and funcs[0].count("DFMA") > 500
    for name in ("_ZN7celeste16unit_walk_kernelILi1E", "_ZN7celeste16unit_walk_kernelILi2E", "_ZN7celeste18unit_moment_kernelE"):
        funcs = [f for f in sass.split("Function : ") if f.startswith(name)]
        assert len(funcs) == 1 and funcs[0].count("DFMA") > 200, name
    # the value / gradient walks stream their pixel records with cp.async (LDGSTS); the Hessian walk does not
    for name, want in (("_ZN7celeste16unit_walk_kernelILi1E", True), ("_ZN7celeste16unit_walk_kernelILi2E", False)):
        f = [f for f in sass.split("Function : ") if f.startswith(name)][0]
        assert ("LDGSTS" in f) == want, name


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_no_device_fails_loudly():
    ndev = C.c_int(-1)
    st = _lib.load().celeste_init(-1, C.byref(ndev))
    assert st == _lib.CELESTE_ERR_NO_DEVICE and ndev.value == 0
    from celeste_jl_b200 import synthetic
    images, patches, vp, _ = synthetic.gen_sample_star_dataset(bands=(3,), H=20, W=20)
    ea = cj.ElboArgs(images, patches, [1], include_kl=False)
    with pytest.raises(_lib.CelesteError) as ei:
        cj.elbo_likelihood(ea, vp)
    assert ei.value.status == _lib.CELESTE_ERR_NO_DEVICE
