"""ctypes binding of libdiffrax_b200.so (the C ABI in include/diffrax_b200.h).

There is deliberately no fallback: if the CUDA library has not been built, or no CUDA
device is present when a solve is requested, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdiffrax_b200.so")
if os.environ.get("DFX_LIB"):  # kernel experiments (tools/build_variant.py): an explicitly named build of the same library
    LIB_PATH = os.path.abspath(os.environ["DFX_LIB"])

ABI_VERSION = 2
F64, F32 = 0, 1
CTRL_CONSTANT, CTRL_PID = 0, 1
LEVY_NONE, LEVY_BI, LEVY_STLA = 0, 1, 2
EVENT_NONE, EVENT_AFFINE, EVENT_STEADY_STATE, EVENT_USER = 0, 1, 2, 3
MAX_EVENTS = 4
SOLVER_IDS = {"tsit5": 0, "dopri5": 1, "dopri8": 2, "heun": 3, "bosh3": 4, "midpoint": 5,
              "ralston": 6, "euler": 7, "shark": 8}
HALF_SOLVER = 0x100  # DFX_HALF_SOLVER: HalfSolver(inner) = HALF_SOLVER | inner id
FIELD_IDS = {"decay": 0, "lotka_volterra": 1, "lorenz": 2, "cr3bp": 3, "mlp": 4, "ou": 5,
             "forced_osc": 6, "vdp": 7, "gbm": 8}
FIELD_OU_MATRIX = 16   # + m: OU drift with a constant [d, m] diffusion matrix
FIELD_USER = 1000      # user functors (fields.CudaField) register ids >= this with dfx_register_launcher


class SolveDesc(C.Structure):
    """struct dfx_solve_desc"""
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("field_id", C.c_int32), ("dim", C.c_int32), ("dtype", C.c_int32), ("solver_id", C.c_int32),
        ("field_params", C.c_void_p), ("n_field_params", C.c_int32),
        ("field_weights", C.c_void_p), ("n_field_weights", C.c_int64),
        ("n_traj", C.c_int64), ("y0", C.c_void_p),
        ("t0", C.c_double), ("t1", C.c_double),
        ("t0_per_traj", C.c_void_p), ("t1_per_traj", C.c_void_p),
        ("dt0", C.c_double),
        ("controller", C.c_int32),
        ("rtol", C.c_double), ("atol", C.c_double), ("pcoeff", C.c_double), ("icoeff", C.c_double),
        ("dcoeff", C.c_double), ("safety", C.c_double), ("factormin", C.c_double), ("factormax", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double), ("force_dtmin", C.c_int32),
        ("error_order", C.c_double),
        ("hairer_initial_step", C.c_int32),
        ("step_ts", C.c_void_p), ("n_step_ts", C.c_int32), ("jump_ts", C.c_void_p), ("n_jump_ts", C.c_int32),
        ("store_rejected_steps", C.c_int32),
        ("save_t0", C.c_int32), ("save_t1", C.c_int32), ("save_steps", C.c_int32), ("save_dense", C.c_int32),
        ("save_ts", C.c_void_p), ("n_save_ts", C.c_int32), ("max_steps", C.c_int32),
        ("ts_out", C.c_void_p), ("ys_out", C.c_void_p), ("stats", C.c_void_p), ("result", C.c_void_p),
        ("save_count", C.c_void_p),
        ("dense_ts", C.c_void_p), ("dense_y0", C.c_void_p), ("dense_y1", C.c_void_p), ("dense_k", C.c_void_p),
        ("dense_count", C.c_void_p),
        ("y_final", C.c_void_p), ("t_final", C.c_void_p),
        ("levy_area", C.c_int32), ("bm_keys", C.c_void_p),
        ("bm_t0", C.c_double), ("bm_t1", C.c_double), ("bm_tol", C.c_double),
        ("threefry_partitionable", C.c_int32), ("bm_dim", C.c_int32),
        ("n_events", C.c_int32), ("event_kind", C.c_int32 * 4), ("event_direction", C.c_int32 * 4),
        ("event_root_find", C.c_int32),
        ("event_params", C.c_void_p), ("n_event_params", C.c_int32),
        ("event_rtol", C.c_double), ("event_atol", C.c_double),
        ("state_in", C.c_void_p), ("state_in_flags", C.c_int32), ("state_out", C.c_void_p),
        ("y_final_device", C.c_void_p), ("t_final_device", C.c_void_p),
        ("totals", C.c_void_p), ("totals_device", C.c_void_p),
        ("dense_lazy_padding", C.c_int32),
        ("n_peers", C.c_int32), ("peer_row_offset", C.c_int64),
        ("peer_y_final", C.c_void_p * 8), ("peer_t_final", C.c_void_p * 8),
        ("traj_args", C.c_void_p), ("n_traj_args", C.c_int32),
    ]


#: every symbol include/diffrax_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "dfx_abi_version", "dfx_last_error", "dfx_device_count", "dfx_num_stages", "dfx_solver_order",
    "dfx_field_dim", "dfx_has_kernel", "dfx_out_size", "dfx_ensemble_solve", "dfx_ensemble_solve_host",
    "dfx_vbt_evaluate", "dfx_broadcast_device_scalar", "dfx_threefry2x32", "dfx_random_split", "dfx_random_normal", "dfx_dense_evaluate", "dfx_dense_derivative", "dfx_dense_pad", "dfx_peer_alloc", "dfx_peer_open", "dfx_peer_close", "dfx_peer_free",
    "dfx_measure_fma_peak", "dfx_measure_int_peak", "dfx_launch_count", "dfx_reset_launch_count",
    "dfx_register_launcher",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


_plugins = {}


def load_plugin(path):
    """dlopen a shared object of user functors built against libdiffrax_b200.so: its static
    registrars call dfx_register_launcher on load."""
    lib()
    path = os.path.abspath(path)
    if path not in _plugins:
        _plugins[path] = C.CDLL(path, mode=C.RTLD_GLOBAL)
    return _plugins[path]


def lib():
    """Load the CUDA library.  Raises LibraryMissing when it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m diffrax_b200.build` "
            "(there is no CPU or PyTorch fallback for the ensemble kernels)")
    # RTLD_GLOBAL: plugins of user functors (load_plugin) bind dfx::register_builtin & co. to THIS loaded instance by name - not
    # through a DT_NEEDED entry, which would map a second copy of the library if the file was rebuilt after it was loaded
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.dfx_abi_version.restype = C.c_int
    L.dfx_last_error.restype = C.c_char_p
    L.dfx_device_count.restype = C.c_int
    L.dfx_num_stages.argtypes = [C.c_int]
    L.dfx_solver_order.argtypes = [C.c_int]
    L.dfx_field_dim.argtypes = [C.c_int]
    L.dfx_has_kernel.argtypes = [C.c_int] * 5
    L.dfx_out_size.argtypes = [C.POINTER(SolveDesc)]
    L.dfx_ensemble_solve.argtypes = [C.POINTER(SolveDesc), C.c_void_p]
    L.dfx_ensemble_solve_host.argtypes = [C.POINTER(SolveDesc), C.c_int]
    L.dfx_vbt_evaluate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_double,
                                   C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.dfx_broadcast_device_scalar.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dfx_threefry2x32.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dfx_random_split.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.dfx_random_normal.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.dfx_dense_evaluate.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]
    L.dfx_dense_derivative.argtypes = L.dfx_dense_evaluate.argtypes
    L.dfx_dense_pad.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p]
    L.dfx_peer_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]
    L.dfx_peer_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.dfx_peer_close.argtypes = [C.c_void_p]
    L.dfx_peer_free.argtypes = [C.c_void_p]
    L.dfx_measure_fma_peak.argtypes = [C.c_int, C.c_int]
    L.dfx_measure_fma_peak.restype = C.c_double
    L.dfx_measure_int_peak.argtypes = [C.c_int]
    L.dfx_measure_int_peak.restype = C.c_double
    L.dfx_launch_count.restype = C.c_int64
    L.dfx_reset_launch_count.restype = None
    if L.dfx_abi_version() != ABI_VERSION:
        raise RuntimeError("libdiffrax_b200.so ABI version mismatch; rebuild")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        msg = lib().dfx_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        if rc == -2:
            raise NotImplementedError(msg)
        raise RuntimeError(f"libdiffrax_b200 error {rc}: {msg}")


def new_desc() -> SolveDesc:
    d = SolveDesc()
    d.struct_size = C.sizeof(SolveDesc)
    d.abi_version = ABI_VERSION
    return d
