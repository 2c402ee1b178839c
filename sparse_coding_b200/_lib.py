"""ctypes binding of libsce.so (include/sce.h). No torch types cross this boundary: only integers, floats and raw
device pointers (``tensor.data_ptr()``).

The library is built in-tree (``make`` / ``__graft_entry__.build()``) next to this file. There is NO fallback:
if it is missing, or the process has no sm_100 device when a plan is created, the engine raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCE_LIB: an alternative build of the same library (kernel A/B experiments, tools/ab_variants.sh)
LIB_PATH = os.environ.get("SCE_LIB") or os.path.join(_HERE, "libsce.so")

SCE_TIED, SCE_UNTIED, SCE_TOPK = 0, 1, 2
SCE_ADAM_FROZEN_T1, SCE_ADAM_STANDARD = 0, 1
SCE_LOSS_COLS = 4
SCE_ARITH_AUTO, SCE_ARITH_BF16X3, SCE_ARITH_F16F8 = 0, 1, 2
ARITH_CODE = {"auto": SCE_ARITH_AUTO, "bf16x3": SCE_ARITH_BF16X3, "f16f8": SCE_ARITH_F16F8}
ARITH_NAME = {SCE_ARITH_BF16X3: "bf16x3", SCE_ARITH_F16F8: "f16f8"}

# every symbol include/sce.h declares (tests check that the built library exports all of them)
EXPORTS = [
    "sce_version", "sce_last_error", "sce_workspace_bytes", "sce_plan_create", "sce_plan_destroy", "sce_prepare",
    "sce_step", "sce_step_host", "sce_forward", "sce_read_code", "sce_grads", "sce_gather_rows",
    "sce_last_launch_count", "sce_get_step_count", "sce_set_step_count", "sce_profile_begin", "sce_profile_end",
    "sce_plan_arith", "sce_input_absmax", "sce_health", "sce_clear_health", "sce_active_counts",
]
PHASES = ["split", "encode", "decode", "losses", "dcode", "dw", "adam"]


class SceDesc(C.Structure):
    _fields_ = [
        ("variant", C.c_int), ("n_models", C.c_int), ("d", C.c_int), ("n", C.c_int), ("batch_max", C.c_int),
        ("x_per_model", C.c_int),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("eps_root", C.c_float),
        ("adam_count_mode", C.c_int), ("fwd_passes", C.c_int), ("bwd_passes", C.c_int), ("norm_floor", C.c_float),
        ("arith", C.c_int), ("topk_k_max", C.c_int), ("centering", C.c_int),
    ]


class SceBuffers(C.Structure):
    _fields_ = [
        ("encoder", C.c_void_p), ("encoder_bias", C.c_void_p), ("decoder", C.c_void_p),
        ("encoder_m", C.c_void_p), ("encoder_v", C.c_void_p), ("bias_m", C.c_void_p), ("bias_v", C.c_void_p),
        ("decoder_m", C.c_void_p), ("decoder_v", C.c_void_p),
        ("l1_alpha", C.c_void_p), ("bias_decay", C.c_void_p), ("coef_mask", C.c_void_p), ("sparsity", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("center_trans", C.c_void_p), ("center_rot", C.c_void_p), ("center_scale", C.c_void_p),
    ]


class SceError(RuntimeError):
    pass


_lib = None


def load():
    """Load libsce.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SceError(
            f"{LIB_PATH} not found: the CUDA engine has not been built. Run `make` at the repository root "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
    lib.sce_version.restype = i
    lib.sce_last_error.restype = C.c_char_p
    lib.sce_workspace_bytes.restype = C.c_size_t
    lib.sce_workspace_bytes.argtypes = [C.POINTER(SceDesc)]
    lib.sce_plan_create.argtypes = [C.POINTER(SceDesc), C.POINTER(SceBuffers), C.POINTER(vp)]
    lib.sce_plan_destroy.argtypes = [vp]
    lib.sce_prepare.argtypes = [vp, vp]
    lib.sce_step.argtypes = [vp, vp, i, vp, vp, vp]
    lib.sce_step_host.argtypes = [vp, vp, i, vp, vp, vp]
    lib.sce_forward.argtypes = [vp, vp, i, vp, vp, vp, vp]
    lib.sce_read_code.argtypes = [vp, i, vp, vp]
    lib.sce_grads.argtypes = [vp, vp, i, vp, vp, vp, vp, vp, vp]
    lib.sce_gather_rows.argtypes = [vp, i, ll, i, vp, i, vp, vp, vp]
    lib.sce_last_launch_count.argtypes = [vp]
    lib.sce_get_step_count.argtypes = [vp]
    lib.sce_get_step_count.restype = ll
    lib.sce_set_step_count.argtypes = [vp, ll]
    lib.sce_profile_begin.argtypes = [vp]
    lib.sce_profile_end.argtypes = [vp, vp, vp]
    lib.sce_plan_arith.argtypes = [vp]
    lib.sce_input_absmax.argtypes = [vp, vp, vp]
    lib.sce_health.argtypes = [vp, vp, vp, vp]
    lib.sce_clear_health.argtypes = [vp, vp]
    lib.sce_active_counts.argtypes = [vp, i, vp, vp]
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError here means header and library disagree
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sce_last_error().decode("utf-8", "replace")
        raise SceError(f"{what} failed (status {rc}): {msg}")
