"""ctypes binding of libcre_b200.so (include/cre.h).

There is deliberately no fallback: if the shared library is missing the import fails loudly, and
every compute entry point raises :class:`CreError` when the C side reports an error (including
"device is not sm_100").  Nothing here computes anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libcre_b200.so"

TOPK_MAX = 8        # candidates one scan pass keeps per query (CRE_TOPK_MAX)
TOPK_LIMIT = 256    # largest k of cre_gallery_topk / cre_merge_topk (CRE_TOPK_LIMIT)

# enum cre_weight_kind
W_PATCH, B_PATCH, PREFIX, LN_F_G, LN_F_B = 0, 1, 2, 3, 4
LN1_G, LN1_B, W_QKV, B_QKV, W_O, B_O, LS1 = 5, 6, 7, 8, 9, 10, 11
LN2_G, LN2_B, W_UP, B_UP, W_DOWN, B_DOWN, LS2 = 12, 13, 14, 15, 16, 17, 18
WEIGHT_KINDS = 19
BF16_KINDS = {W_PATCH, W_QKV, W_O, W_UP, W_DOWN}

# enum cre_kernel_id
KERNEL_NAMES = ["preprocess", "fill_prefix", "gemm_patch", "layernorm", "gemm_qkv", "attention", "gemm_resid", "gemm_gelu",
                "final_norm_mean", "pool_clips", "split_hi_lo", "fill_topk", "gemm_topk", "merge_topk", "gemm_plain",
                "gallery_update", "row_stats", "fold_ln_weights", "roi_tables", "attention_exact", "gemm_resid_mlp"]
# kernels whose `work` is FLOPs (tensor-bound); the others report bytes (HBM-bound)
FLOP_KERNELS = {"gemm_patch", "gemm_qkv", "attention", "gemm_resid", "gemm_resid_mlp", "gemm_gelu", "gemm_plain"}

# enum cre_gemm_epilogue
EPI_BF16, EPI_F32, EPI_GELU, EPI_RESID, EPI_NONE, EPI_RESID_LN, EPI_RESID_LN3, EPI_RESID_SP, EPI_RESID_SP3 = 0, 1, 3, 4, 7, 8, 9, 10, 11


class CreError(RuntimeError):
    """Raised when a libcre_b200 call returns a non-zero status."""


class ModelCfg(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32),
        ("layers", C.c_int32),
        ("heads", C.c_int32),
        ("mlp", C.c_int32),
        ("patch", C.c_int32),
        ("registers", C.c_int32),
        ("rope_theta", C.c_float),
        ("ln_eps", C.c_float),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_cfgp = C.POINTER(ModelCfg)
_f3 = C.POINTER(C.c_float)

# name -> (restype, argtypes); must list every symbol include/cre.h declares (tests check this)
PROTOTYPES = {
    "cre_last_error": (C.c_char_p, []),
    "cre_abi_version": (_i32, []),
    "cre_packed_weights_bytes": (_i64, [_cfgp]),
    "cre_weight_offset": (_i64, [_cfgp, _i32, _i32]),
    "cre_weight_elems": (_i64, [_cfgp, _i32, _i32]),
    "cre_create": (_i32, [_cfgp, _vp, _i32, C.POINTER(_vp)]),
    "cre_destroy": (_i32, [_vp]),
    "cre_workspace_bytes": (_i64, [_cfgp, _i32, _i32, _i32]),
    "cre_preprocess_patchify": (_i32, [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _i32, _i32, _i32, _f3, _f3, _vp, _vp]),
    "cre_roi_scratch_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32]),
    "cre_preprocess_patchify_roi": (_i32, [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _i32, _vp, _i32, _i32, _i32, _f3, _f3, _vp, _i64, _vp, _vp]),
    "cre_vit_forward": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "cre_pool_clips": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "cre_gallery_scratch_bytes": (_i64, [_i32, _i32, _i32]),
    "cre_gallery_topk": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "cre_merge_topk": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cre_gallery_update_row": (_i32, [_vp, _vp, _i32, _i32, _vp, _f32, _vp]),
    "cre_gemm_bf16": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp]),
    "cre_layernorm_bf16": (_i32, [_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp]),
    "cre_row_stats": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "cre_row_stats_split": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cre_fold_ln_weights": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "cre_gemm_ln": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _vp, _vp, _i32, _vp]),
    "cre_attention": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "cre_attention_scratch_bytes": (_i64, [_i32, _i32]),
    "cre_set_cta_group": (_i32, [_i32]),
    "cre_set_tuning": (_i32, [C.c_char_p, _i32]),
    "cre_kernel_launches": (_i64, []),
    "cre_profile_start": (_i32, [_i32]),
    "cre_profile_stop": (_i32, [_vp, _vp, _vp, _i32]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the library (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("CRE_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C vision_sam3_yolo_lameless_b200/csrc`). There is no CPU/PyTorch fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().cre_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str) -> None:
    if status != 0:
        raise CreError(f"{what} failed (status {status}): {last_error()}")


def check_size(value: int, what: str) -> int:
    if value < 0:
        raise CreError(f"{what} failed: {last_error()}")
    return int(value)


def kernel_launches() -> int:
    """Kernels launched by libcre_b200 in this process so far."""
    return int(load().cre_kernel_launches())


def profile_start(max_launches: int = 1 << 16) -> None:
    check(load().cre_profile_start(int(max_launches)), "cre_profile_start")


def profile_stop(cap: int = 1 << 16):
    """-> list of (kernel name, milliseconds, work) for every launch since profile_start (synchronises)."""
    ids = (C.c_int32 * cap)()
    ms = (C.c_float * cap)()
    work = (C.c_double * cap)()
    n = load().cre_profile_stop(ids, ms, work, cap)
    if n < 0:
        raise CreError(f"cre_profile_stop failed: {last_error()}")
    return [(KERNEL_NAMES[ids[i]], float(ms[i]), float(work[i])) for i in range(n)]


def set_tuning(key: str, value: int) -> None:
    check(load().cre_set_tuning(key.encode(), int(value)), f"cre_set_tuning({key})")
