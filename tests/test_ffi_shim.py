"""The jax.ffi shim (diffrax_b200/csrc/ffi/xla_ffi_shim.cc) is compiled and linked on every run - against jaxlib's
xla/ffi/api/ffi.h when jax is importable, else against the compile-check stub in tests/ffi_stub, whose Binding::To
static_asserts that the bound operand / result / attribute list matches the handler's signature.  (No compute calls.)"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_links_and_exports_handlers(tmp_path, cuda_lib):
    from diffrax_b200 import build
    # next to libdiffrax_b200.so, where its rpath ($ORIGIN) finds the core library
    out, kind = build.build_ffi(out=os.path.join(build.LIBDIR, "libdfx_xla_ffi_check.so"), verbose=False)
    assert kind in ("jaxlib", "stub")
    lib = ctypes.CDLL(out)
    for name in ("DfxEnsembleSolveF64", "DfxEnsembleSolveF32", "DfxDenseEvaluateF64", "DfxDenseEvaluateF32"):
        assert hasattr(lib, name), name


def test_shim_binds_every_descriptor_block():
    """The handler wires the blocks the first version dropped: per-trajectory t0/t1, bm_dim, dense outputs, step_ts /
    jump_ts, events and resumed state - and is registered for batched (not per-trajectory) launches."""
    src = open(os.path.join(ROOT, "diffrax_b200", "csrc", "ffi", "xla_ffi_shim.cc")).read()
    for needle in ("t0_per_traj", "t1_per_traj", "bm_dim", "dense_ts", "dense_k", "step_ts", "jump_ts", "event_kind",
                   "state_in", "state_out", "store_rejected_steps", "y_final", "t_final", "field_weights"):
        assert f"d.{needle}" in src, needle
    assert "(void)" not in src
    py = open(os.path.join(ROOT, "diffrax_b200", "jax_ffi.py")).read()
    assert 'vmap_method="expand_dims"' in py
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert 'vmap_method="sequential"' not in doc.split("dense_evaluate")[0]


def test_stub_rejects_a_mismatched_binding(tmp_path):
    """The stub is a real check: an implementation whose parameter list does not match the binding fails to compile."""
    from diffrax_b200 import build
    inc, kind = build.ffi_include_dir()
    src = tmp_path / "bad.cc"
    src.write_text('''
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;
static ffi::Error Impl(ffi::Buffer<ffi::F64> a, ffi::ResultBuffer<ffi::F64> out, int32_t attr) { return ffi::Error::Success(); }
XLA_FFI_DEFINE_HANDLER_SYMBOL(Bad, Impl, ffi::Ffi::Bind().Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Attr<double>("attr").Attr<int32_t>("extra"));
''')
    r = subprocess.run([os.environ.get("CXX", "g++"), "-std=c++17", "-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
    assert r.returncode != 0
    good = tmp_path / "good.cc"
    good.write_text('''
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;
static ffi::Error Impl(ffi::Buffer<ffi::F64> a, ffi::ResultBuffer<ffi::F64> out, int32_t attr) { return ffi::Error::Success(); }
XLA_FFI_DEFINE_HANDLER_SYMBOL(Good, Impl, ffi::Ffi::Bind().Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Attr<int32_t>("attr"));
''')
    r = subprocess.run([os.environ.get("CXX", "g++"), "-std=c++17", "-fsyntax-only", "-I", inc, str(good)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_jax_binding_module_is_syntactically_valid():
    import ast
    ast.parse(open(os.path.join(ROOT, "diffrax_b200", "jax_ffi.py")).read())
    try:
        import jax  # noqa: F401
    except Exception:  # noqa: BLE001
        import pytest
        with pytest.raises(ImportError):
            import diffrax_b200.jax_ffi  # noqa: F401


def test_python_binding_matches_the_handler_signature():
    """diffrax_b200/jax_ffi.py passes exactly the attributes the C++ binding declares (same names), ten operands and twelve
    results for the solve, six operands and one result for the dense evaluation."""
    import re
    src = open(os.path.join(ROOT, "diffrax_b200", "csrc", "ffi", "xla_ffi_shim.cc")).read()
    py = open(os.path.join(ROOT, "diffrax_b200", "jax_ffi.py")).read()
    solve_bind = src[src.index("#define DFX_BIND_SOLVE"):src.index("#define DFX_BIND_DENSE")]
    dense_bind = src[src.index("#define DFX_BIND_DENSE"):src.index("XLA_FFI_DEFINE_HANDLER_SYMBOL(DfxEnsembleSolveF64")]
    attrs = re.findall(r'\.Attr<[^>]+>+\("(\w+)"\)', solve_bind)
    call = py[py.index("outs = call("):py.index("ts_o, ys_o, stats")]
    kwargs = re.findall(r"\b(\w+)=", call)
    assert attrs and sorted(set(kwargs)) == sorted(attrs), (set(attrs) ^ set(kwargs))
    assert solve_bind.count(".Arg<") == 10 and solve_bind.count(".Ret<") == 12
    out_types = py[py.index("out_types = ("):py.index('name = "DfxEnsembleSolveF64"')]
    assert out_types.count("S(") == 12
    dattrs = re.findall(r'\.Attr<[^>]+>+\("(\w+)"\)', dense_bind)
    dcall = py[py.index("return call(dense["):py.index("def lorenz_dopri5_example")]
    assert sorted(re.findall(r"\b(\w+)=", dcall)) == sorted(dattrs)
    assert dense_bind.count(".Arg<") == 6 and dense_bind.count(".Ret<") == 1
