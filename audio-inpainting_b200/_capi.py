"""ctypes signatures of include/ainmf.h (shared by the product loader and the test harness)."""
from __future__ import annotations

import ctypes as C

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_WORKSPACE, ERR_ALL_BAD, ERR_COMM = -1, -2, -3, -4, -5, -6
SOLVER_CD, SOLVER_MU, SOLVER_MU_KL = 0, 1, 2

EXPORTS = [
    "ainmf_create", "ainmf_destroy", "ainmf_last_error", "ainmf_version", "ainmf_params_default",
    "ainmf_stft_geometry", "ainmf_padded_rank", "ainmf_stft", "ainmf_gap_mask", "ainmf_nmf_fit", "ainmf_istft",
    "ainmf_workspace_bytes", "ainmf_inpaint", "ainmf_inpaint_host", "ainmf_load_pcm16", "ainmf_store_pcm16",
    "ainmf_comm_unique_id", "ainmf_comm_init", "ainmf_shard_plan", "ainmf_sharded_workspace_bytes",
    "ainmf_inpaint_sharded", "ainmf_launch_count", "ainmf_profile", "ainmf_comm_set_callbacks",
    "ainmf_find_main_gap", "ainmf_find_gaps", "ainmf_linear_interp", "ainmf_blend_boundaries", "ainmf_snr_db", "ainmf_apply_gaps",
    "ainmf_set_window", "ainmf_standard_normal", "ainmf_comm_transport", "ainmf_inpaint_host_pcm16",
    "ainmf_host_chunk_schedule",
]

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p)
SENDRECV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p)


class Params(C.Structure):
    """struct ainmf_params (include/ainmf.h)."""
    _fields_ = [
        ("batch", C.c_int32), ("n_samples", C.c_int64), ("n_fft", C.c_int32), ("hop", C.c_int32),
        ("rank", C.c_int32), ("max_iter", C.c_int32), ("tol", C.c_float), ("solver", C.c_int32),
        ("seed", C.c_uint32), ("threshold", C.c_float), ("frac_num", C.c_int32), ("frac_den", C.c_int32),
        ("col_start", C.c_int32), ("col_end", C.c_int32), ("n_outer", C.c_int32),
    ]


def bind(lib: C.CDLL) -> C.CDLL:
    vp, i32, i64, f32, u32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32, C.c_size_t
    P = C.POINTER
    sig = {
        "ainmf_create": (C.c_int, [P(vp), C.c_int]),
        "ainmf_destroy": (C.c_int, [vp]),
        "ainmf_last_error": (C.c_char_p, [vp]),
        "ainmf_version": (C.c_char_p, []),
        "ainmf_params_default": (None, [P(Params)]),
        "ainmf_stft_geometry": (C.c_int, [i64, i32, i32, P(i32), P(i32), P(i32)]),
        "ainmf_padded_rank": (i32, [i32]),
        "ainmf_stft": (C.c_int, [vp, vp, i32, i64, i32, i32, vp, vp, vp]),
        "ainmf_gap_mask": (C.c_int, [vp, vp, i32, i64, i32, i32, f32, i32, i32, vp, vp, vp, vp]),
        "ainmf_nmf_fit": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, f32, i32, u32, vp, vp, vp, vp, vp, vp, vp]),
        "ainmf_istft": (C.c_int, [vp, vp, i32, i32, i32, i32, i64, vp, vp]),
        "ainmf_workspace_bytes": (sz, [vp, P(Params)]),
        "ainmf_inpaint": (C.c_int, [vp, P(Params), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
        "ainmf_inpaint_host": (C.c_int, [vp, P(Params), vp, vp, vp, vp, vp, sz]),
        "ainmf_inpaint_host_pcm16": (C.c_int, [vp, P(Params), vp, i32, vp, vp, vp, vp, vp, sz]),
        "ainmf_host_chunk_schedule": (C.c_int, [i64, i64, i32, vp, i32, vp]),
        "ainmf_load_pcm16": (C.c_int, [vp, vp, i32, i64, i32, vp, vp, vp]),
        "ainmf_store_pcm16": (C.c_int, [vp, vp, i64, vp, vp]),
        "ainmf_find_main_gap": (C.c_int, [vp, vp, i32, i64, f32, vp, vp]),
        "ainmf_find_gaps": (C.c_int, [vp, vp, i32, i64, f32, i32, vp, i32, vp, vp]),
        "ainmf_linear_interp": (C.c_int, [vp, vp, i32, i64, f32, vp, vp, vp]),
        "ainmf_apply_gaps": (C.c_int, [vp, vp, i32, i64, vp, vp, i32, vp]),
        "ainmf_blend_boundaries": (C.c_int, [vp, vp, vp, i64, i64, i64, i32, vp, vp]),
        "ainmf_snr_db": (C.c_int, [vp, vp, vp, i64, i64, P(C.c_double), vp]),
        "ainmf_launch_count": (C.c_ulonglong, []),
        "ainmf_profile": (C.c_int, [vp, i32, vp, vp]),
        "ainmf_set_window": (C.c_int, [vp, i32, vp]),
        "ainmf_standard_normal": (C.c_int, [vp, C.c_uint32, i64, vp, vp]),
        "ainmf_comm_unique_id": (C.c_int, [vp]),
        "ainmf_comm_init": (C.c_int, [vp, vp, i32, i32]),
        "ainmf_comm_transport": (C.c_int, [vp]),
        "ainmf_comm_set_callbacks": (C.c_int, [vp, i32, i32, ALLREDUCE_FN, SENDRECV_FN, vp]),
        "ainmf_shard_plan": (C.c_int, [i64, i32, i32, i32, i32, P(i32), P(i32), P(i64), P(i64), P(i64), P(i64)]),
        "ainmf_sharded_workspace_bytes": (sz, [vp, P(Params)]),
        "ainmf_inpaint_sharded": (C.c_int, [vp, P(Params), vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)     # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return lib


class AinmfError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ainmf error {code}: {message}")
        self.code = code


def default_params(lib: C.CDLL, **kw) -> Params:
    p = Params()
    lib.ainmf_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k!r}")
        setattr(p, k, v)
    return p


def stft_geometry(lib: C.CDLL, n_samples: int, n_fft: int, hop: int):
    T, F, ldf = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.ainmf_stft_geometry(n_samples, n_fft, hop, C.byref(T), C.byref(F), C.byref(ldf))
    if rc:
        raise AinmfError(rc, f"invalid STFT geometry N={n_samples} n_fft={n_fft} hop={hop}")
    return T.value, F.value, ldf.value
