"""ctypes binding of include/snsde.h.  Fails loudly: there is no CPU or eager fallback."""
import ctypes
import pathlib

HERE = pathlib.Path(__file__).resolve().parent
import os
LIB_PATH = HERE / ("libsnsde_trace.so" if os.environ.get("SNSDE_TRACE_BUILD") else "libsnsde.so")

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_WEIGHTS, ERR_INTERNAL = 0, -1, -2, -3, -4, -5
ABI_VERSION = 4
FAMILY_BENCHMARK, FAMILY_TUTORIAL_LSDE, FAMILY_LATENT_SDE = 0, 1, 2
METHOD = {"euler": 0, "milstein": 1, "srk": 2}
PRECISION = {"fp32": 0, "tc": 1, "auto": 2}


class ModelDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "family", "input_option", "noise_option", "input_channels", "hidden", "hidden_hidden",
        "num_hidden_layers", "method", "precision")]


class Step(ctypes.Structure):
    _fields_ = [("t0", ctypes.c_float), ("h", ctypes.c_float), ("sqrt_h", ctypes.c_float),
                ("sin_t0", ctypes.c_float), ("cos_t0", ctypes.c_float), ("interval", ctypes.c_int32),
                ("frac", ctypes.c_float), ("emit_begin", ctypes.c_int32), ("emit_end", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class Emit(ctypes.Structure):
    _fields_ = [("slot", ctypes.c_int32), ("w_prev", ctypes.c_float), ("w_curr", ctypes.c_float)]


class Point(ctypes.Structure):
    _fields_ = [("t", ctypes.c_float), ("sin_t", ctypes.c_float), ("cos_t", ctypes.c_float),
                ("frac", ctypes.c_float), ("interval", ctypes.c_int32)]


EXPORTS = ("snsde_plan_fma_variant", "snsde_natural_coeffs_missing", "snsde_initial_state", "snsde_readout_head", "snsde_plan_status_nowait", "snsde_backward", "snsde_backward_workspace_bytes", "snsde_abi_version", "snsde_last_error", "snsde_weight_count", "snsde_plan_create",
           "snsde_plan_destroy", "snsde_plan_set_weights", "snsde_plan_kernel_kind", "snsde_forward",
           "snsde_philox_fill", "snsde_plan_launch_count", "snsde_plan_status", "snsde_hermite_coeffs",
           "snsde_natural_coeffs", "snsde_fill_missing")

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    """Load libsnsde.so (built by ``build.py``).  Raises if it is missing - by design."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EngineError(
            f"{LIB_PATH} not found: build it with `python stable-neural-sdes_b200/build.py` "
            "(or __graft_entry__.build()).  This engine has no CPU/eager fallback.")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
    lib.snsde_abi_version.restype = ctypes.c_int
    lib.snsde_last_error.restype = ctypes.c_char_p
    lib.snsde_weight_count.restype = i64
    lib.snsde_weight_count.argtypes = [ctypes.POINTER(ModelDesc)]
    lib.snsde_plan_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.c_int, ctypes.POINTER(vp)]
    lib.snsde_plan_destroy.argtypes = [vp]
    lib.snsde_plan_set_weights.argtypes = [vp, vp, i64, ctypes.c_int, vp]
    lib.snsde_plan_kernel_kind.argtypes = [vp]
    lib.snsde_plan_launch_count.argtypes = [vp]
    lib.snsde_plan_fma_variant.argtypes = [vp]
    lib.snsde_plan_launch_count.restype = i64
    lib.snsde_plan_status.argtypes = [vp, vp]
    lib.snsde_hermite_coeffs.argtypes = [vp, vp, i32, i32, i32, vp, ctypes.c_int, vp]
    lib.snsde_natural_coeffs.argtypes = [vp, vp, i32, i32, i32, vp, vp, ctypes.c_int, vp]
    lib.snsde_fill_missing.argtypes = [vp, vp, i32, i32, i32, vp, ctypes.c_int, vp]
    lib.snsde_natural_coeffs_missing.argtypes = [vp, vp, i32, i32, i32, vp, ctypes.c_int, vp]
    lib.snsde_plan_status_nowait.argtypes = [vp]
    lib.snsde_initial_state.argtypes = [vp, i64, i32, i32, i32, i32, ctypes.c_float, vp, vp, i32, vp, ctypes.c_int, vp]
    lib.snsde_readout_head.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, ctypes.c_int, vp]
    lib.snsde_forward.argtypes = [vp, vp, i64, i32, vp, i32, vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, u64, u64, vp, vp]
    lib.snsde_philox_fill.argtypes = [u64, u64, i32, i32, i32, vp, vp, vp, ctypes.c_int, vp]
    lib.snsde_backward_workspace_bytes.restype = i64
    lib.snsde_backward_workspace_bytes.argtypes = [vp, i32, i32]
    lib.snsde_backward.argtypes = [vp, vp, i64, i32, i32, vp, i32, vp, vp, vp, vp, vp, u64, u64, vp, vp, vp, i64, vp]
    for name in EXPORTS:
        getattr(lib, name)
    if lib.snsde_abi_version() != ABI_VERSION:
        raise EngineError("libsnsde.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        msg = load().snsde_last_error().decode()
        if rc in (ERR_BAD_ARG, ERR_UNSUPPORTED):
            raise ValueError(f"snsde: {msg}")          # torchsde raises ValueError for contract violations
        raise EngineError(f"snsde (status {rc}): {msg}")
    return rc
