"""ctypes binding of librwmpt.so (include/rwmpt.h).  No CPU fallback: if the library is missing and cannot
be built, or no CUDA device is present when a kernel is requested, the call raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RWMPT_LIB", os.path.join(_PKG, "librwmpt.so"))  # RWMPT_LIB: A/B-test a variant build

# enums of include/rwmpt.h
T_ROUGH_CARPET, T_THREE_MIXTURE, T_FULL_ROSENBROCK, T_EVEN_ROSENBROCK, T_HYBRID_ROSENBROCK = 0, 1, 2, 3, 4
T_NEAL_FUNNEL, T_HYPERCUBE, T_IID_GAMMA, T_IID_BETA, T_SCALED_MVN, T_MVN_DIAG = 5, 6, 7, 8, 9, 10
T_MVN_DENSE, T_SUPER_FUNNEL = 11, 12
PARAM_HEADER = 16
P_NORMAL, P_LAPLACE, P_UNIFORM_RADIUS = 0, 1, 2
SWAP_REFERENCE, SWAP_EXCHANGE = 0, 1
MATH_FAST, MATH_IEEE = 0, 1
STORE_NONE, STORE_COLD, STORE_ALL = 0, 1, 2
EINVAL, ENOTSUP, ECUDA = -1, -2, -3

SWAP_MODES = {"reference": SWAP_REFERENCE, "exchange": SWAP_EXCHANGE}
MATH_MODES = {"fast": MATH_FAST, "ieee": MATH_IEEE}
STORE_MODES = {"none": STORE_NONE, "cold": STORE_COLD, "all": STORE_ALL}


class TargetT(C.Structure):
    _fields_ = [("family", C.c_int32), ("dim", C.c_int32), ("params", C.c_void_p), ("n_params", C.c_int64)]


class RunArgs(C.Structure):
    """Mirror of rwmpt_run_args_t -- field order and types must match include/rwmpt.h exactly
    (rwmpt_sizeof_run_args() is checked at load time)."""
    _fields_ = [
        ("target", TargetT),
        ("proposal_family", C.c_int32), ("n_temps", C.c_int32),
        ("prop_scale", C.c_void_p), ("prop_dim_scale", C.c_void_p), ("beta", C.c_void_p),
        ("n_ladders", C.c_int64), ("n_steps", C.c_int64), ("burn_in", C.c_int64), ("step_offset", C.c_int64),
        ("swap_every", C.c_int32), ("swap_mode", C.c_int32),
        ("state", C.c_void_p), ("logp", C.c_void_p),
        ("seed", C.c_uint64), ("chain_id_base", C.c_int64),
        ("samples", C.c_void_p), ("sample_logp", C.c_void_p),
        ("store_mode", C.c_int32), ("math_mode", C.c_int32),
        ("store_start", C.c_int64), ("thin", C.c_int64), ("sample_stride", C.c_int64), ("sample_rows", C.c_int64),
        ("accept_count", C.c_void_p), ("sq_jump_sum", C.c_void_p),
        ("swap_accepts", C.c_void_p), ("swap_last_attempt", C.c_void_p),
        ("inj_increments", C.c_void_p), ("inj_uniforms", C.c_void_p), ("inj_swap_uniforms", C.c_void_p),
        ("decisions", C.c_void_p), ("swap_decisions", C.c_void_p),
        ("lanes_per_chain", C.c_int32), ("schedule", C.c_int32),
    ]


_lib = None
_lock = threading.Lock()


def _declare(lib):
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    lib.rwmpt_version.restype = C.c_int
    lib.rwmpt_last_error.restype = C.c_char_p
    lib.rwmpt_sizeof_run_args.restype = u64
    lib.rwmpt_rwm_run.argtypes = [C.POINTER(RunArgs), vp]
    lib.rwmpt_pt_run.argtypes = [C.POINTER(RunArgs), vp]
    lib.rwmpt_count_swap_rounds.argtypes = [i64, i64, i64, i32]
    lib.rwmpt_count_swap_rounds.restype = i64
    lib.rwmpt_pick_lanes.argtypes = [i32, i32, i64, i32, C.POINTER(i32)]
    lib.rwmpt_pick_geometry.argtypes = [C.POINTER(RunArgs), C.POINTER(i32), C.POINTER(i32)]
    lib.rwmpt_log_density.argtypes = [C.POINTER(TargetT), vp, i64, vp, i32, vp]
    lib.rwmpt_proposal_sample.argtypes = [i32, i32, C.c_float, vp, i64, u64, i64, vp, vp]
    lib.rwmpt_swap_prob_estimate.argtypes = [C.POINTER(TargetT), C.c_float, C.c_float, i64, u64, i64, vp, vp]
    lib.rwmpt_pt_swap.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp, u64, i64, i64, vp, vp, vp]
    lib.rwmpt_esjd_reduce.argtypes = [vp, i64, i64, i64, i64, i32, vp, vp, vp]
    lib.rwmpt_debug_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.rwmpt_probe_peaks.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.rwmpt_probe_peaks.restype = C.c_int
    lib.rwmpt_probe_issue.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.rwmpt_probe_issue.restype = C.c_int
    lib.rwmpt_run_host.argtypes = [C.POINTER(RunArgs), i32, C.POINTER(u64), C.POINTER(u64)]
    for name in ("rwmpt_rwm_run", "rwmpt_pt_run", "rwmpt_pick_lanes", "rwmpt_pick_geometry", "rwmpt_log_density", "rwmpt_proposal_sample",
                 "rwmpt_pt_swap", "rwmpt_esjd_reduce", "rwmpt_debug_philox", "rwmpt_run_host", "rwmpt_swap_prob_estimate"):
        getattr(lib, name).restype = C.c_int


def load():
    """Load (building in-tree first if absent) librwmpt.so.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                import importlib
                importlib.import_module(__package__ + ".build").build()
            lib = C.CDLL(LIB_PATH)
            _declare(lib)
            if lib.rwmpt_sizeof_run_args() != C.sizeof(RunArgs):
                raise RuntimeError(f"ABI mismatch: rwmpt_run_args_t is {lib.rwmpt_sizeof_run_args()} bytes in "
                                   f"librwmpt.so but {C.sizeof(RunArgs)} in the Python binding")
            _lib = lib
    return _lib


def check(rc: int):
    if rc >= 0:
        return rc
    msg = load().rwmpt_last_error().decode("utf-8", "replace")
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOTSUP:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def require_cuda(device) -> torch.device:
    """The sampling path exists only as sm_100a CUDA: there is deliberately no CPU fallback."""
    dev = torch.device(device) if device is not None else torch.device("cuda")
    if dev.type != "cuda":
        raise RuntimeError(f"rwm_pt_pytorch_b200 runs only on CUDA devices (got device='{dev}'); "
                           "there is no CPU fallback for the sampling hot path")
    if not torch.cuda.is_available():
        raise RuntimeError("rwm_pt_pytorch_b200 needs a CUDA device (none visible); there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def target_struct(family: int, dim: int, params: torch.Tensor) -> TargetT:
    return TargetT(family, dim, params.data_ptr(), params.numel())


def exported_symbols():
    """Names declared in include/rwmpt.h (used by the CPU test that checks the library exports them all)."""
    return ["rwmpt_version", "rwmpt_last_error", "rwmpt_sizeof_run_args", "rwmpt_rwm_run", "rwmpt_pt_run",
            "rwmpt_count_swap_rounds", "rwmpt_pick_lanes", "rwmpt_pick_geometry", "rwmpt_log_density", "rwmpt_proposal_sample",
            "rwmpt_pt_swap", "rwmpt_esjd_reduce", "rwmpt_probe_peaks", "rwmpt_probe_issue", "rwmpt_debug_philox", "rwmpt_run_host",
            "rwmpt_swap_prob_estimate"]
