"""Build libdiffrax_b200.so in-tree with nvcc for sm_100a (no torch extension machinery).

    python -m diffrax_b200.build [--force] [-j N]

Each translation unit under csrc/ is compiled to an object in csrc/_obj/ (in parallel) and
linked into diffrax_b200/lib/libdiffrax_b200.so.  The .so travels to the GPU box with the
gpurun snapshot; it is git-ignored.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdiffrax_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v", "-I", os.path.join(HERE, "..", "include")] + os.environ.get("DFX_NVCC_EXTRA", "").split()


# headers that only generated plugin sources include (fields.CudaField(wide=True)): not dependencies of the library's own objects
PLUGIN_ONLY_HEADERS = ("wide_kernel.cuh",)


def _headers(plugin=False):
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "diffrax_b200.h")]
    return hs if plugin else [h for h in hs if os.path.basename(h) not in PLUGIN_ONLY_HEADERS]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    log = obj + ".log"
    t0 = time.time()
    r = subprocess.run([NVCC, *ARCH, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    return src, obj, r.returncode, time.time() - t0, r.stderr


def build(force=False, jobs=None, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = _headers()
    todo = [s for s in srcs if force or _stale(os.path.join(OBJ, os.path.basename(s)[:-3] + ".o"), [s] + hdrs)]
    jobs = jobs or min(len(todo) or 1, os.cpu_count() or 4)
    if todo:
        with cf.ThreadPoolExecutor(jobs) as ex:
            for src, obj, rc, dt, err in ex.map(_compile, todo):
                if verbose:
                    print(f"[build] {os.path.basename(src)} rc={rc} {dt:.1f}s", flush=True)
                if rc != 0:
                    sys.stderr.write(err[-6000:])
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or todo or _stale(LIB, objs):
        r = subprocess.run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            raise RuntimeError("link failed")
        if verbose:
            print(f"[build] linked {LIB}", flush=True)
    return LIB


USER_DIR = os.path.join(LIBDIR, "user")


def build_user_field(tag, source, verbose=False):
    """Compile one generated translation unit (fields.CudaField: a user functor + ONE DFX_REGISTER) against csrc/launch.cuh
    into lib/user/dfx_user_<tag>.so (a plugin of the loaded libdiffrax_b200.so).  Cached: the tag carries the content hash of the functor,
    and the object is rebuilt when the kernel headers are newer
    (the plugin binds to the library by name at load time, so relinking the library does not invalidate it)."""
    os.makedirs(USER_DIR, exist_ok=True)
    out = os.path.join(USER_DIR, f"dfx_user_{tag}.so")
    src = os.path.join(USER_DIR, f"dfx_user_{tag}.cu")
    if not os.path.exists(LIB):
        raise RuntimeError(f"{LIB} not found: build it with `python -m diffrax_b200.build` first")
    if os.path.exists(src) and open(src).read() == source and not _stale(out, [src] + _headers(plugin=True)):
        return out
    if not os.path.exists(NVCC):
        raise RuntimeError(f"compiling a CudaField needs nvcc ({NVCC} not found; set NVCC)")
    import threading
    uniq = f".tmp{os.getpid()}_{threading.get_ident()}"   # concurrent threads / ranks may build the same tag
    with open(src + uniq, "w") as f:
        f.write(source)
    os.replace(src + uniq, src)
    tmp = out + uniq
    flags = [x for x in FLAGS if x not in ("-Xptxas", "-v")]
    # (no -ldiffrax_b200: the registrar's symbols stay undefined in the plugin and are resolved at dlopen time against the
    # library instance the process has already loaded RTLD_GLOBAL - see _lib.lib)
    cmd = [NVCC, *ARCH, *flags, "-I", CSRC, "-shared", "-o", tmp, src]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed on the generated CudaField source " + src + ":\n" + (r.stderr or r.stdout)[-4000:])
    os.replace(tmp, out)   # atomic
    if verbose:
        print(f"[build] {out} {time.time() - t0:.1f}s", flush=True)
    return out


FFI_SRC = os.path.join(CSRC, "ffi", "xla_ffi_shim.cc")
FFI_LIB = os.path.join(LIBDIR, "libdfx_xla_ffi.so")
FFI_STUB = os.path.join(HERE, "..", "tests", "ffi_stub")


def ffi_include_dir():
    """(dir, kind): jaxlib's XLA FFI headers when jax is importable, else the compile-check stub under tests/ffi_stub."""
    try:
        import jax.ffi  # noqa: F401
        return jax.ffi.include_dir(), "jaxlib"
    except Exception:  # noqa: BLE001
        return FFI_STUB, "stub"


def build_ffi(out=None, verbose=True):
    """Compile + link csrc/ffi/xla_ffi_shim.cc (the jax.ffi handlers) against libdiffrax_b200.so.
    With jaxlib's headers the result is the loadable handler library; with the stub it is a compile / link check only."""
    build(verbose=False)
    inc, kind = ffi_include_dir()
    out = out or FFI_LIB
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", inc, "-I", os.path.join(HERE, "..", "include"),
           "-I", "/usr/local/cuda/include", FFI_SRC, "-L", LIBDIR, "-ldiffrax_b200", "-Wl,-rpath,$ORIGIN", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-6000:])
        raise RuntimeError("building the jax.ffi shim failed")
    if verbose:
        print(f"[build] {out} (XLA FFI headers: {kind})", flush=True)
    return out, kind


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("--ffi", action="store_true", help="also build the jax.ffi handler library (lib/libdfx_xla_ffi.so)")
    a = ap.parse_args()
    build(a.force, a.j)
    if a.ffi:
        build_ffi()
