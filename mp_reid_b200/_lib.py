"""ctypes binding of libmpreid_b200.so (C ABI: include/mpreid_b200.h).

The library is the product; there is no Python / CPU fallback.  If the shared object is missing it
is built in-tree with nvcc (mp_reid_b200/build.py); if that is impossible the import of the compute
path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmpreid_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "mpreid_b200.h")

# enums of include/mpreid_b200.h
SQEUCLID, ARCCOS, ONE_MINUS_DOT, SQRT_EUCLID, DOT = 0, 1, 2, 3, 4
FP32_SIMT, X3TF32, BF16, X3FP16, X2FP16 = 0, 1, 2, 3, 4
JUNK_NONE, JUNK_PID_CAM = 0, 1

METRICS = {"sqeuclid": SQEUCLID, "euclidean": SQEUCLID, "arccos": ARCCOS, "cosine": ARCCOS,
           "one_minus_dot": ONE_MINUS_DOT, "1-cos": ONE_MINUS_DOT, "sqrt_euclid": SQRT_EUCLID, "dot": DOT}
PRECISIONS = {"simt": FP32_SIMT, "fp32_simt": FP32_SIMT, "3xtf32": X3TF32, "bf16": BF16, "3xfp16": X3FP16, "fp32": X3FP16, "2xfp16": X2FP16}
JUNKS = {"none": JUNK_NONE, "pid_cam": JUNK_PID_CAM}

_p, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

SIGNATURES = {
    "mpreid_last_error": (C.c_char_p, []),
    "mpreid_abi_version": (_i32, []),
    "mpreid_device_info": (_i32, [_i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "mpreid_prep_rows": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "mpreid_dist_matrix": (_i32, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i64, _p, _p]),
    "mpreid_dist_matrix_symmetric": (_i32, [_p, _p, _p, _p, _i64, _i64, _i64, _i32, _i32, _p, _i64, _p, _p]),
    "mpreid_rank_eval_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "mpreid_rank_eval": (_i32, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _sz, _i64, _p, _p]),
    "mpreid_eval_features_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i64, _i32]),
    "mpreid_eval_features": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _p, _i64,
                                    _p, _sz, _i64, _p, _p]),
    "mpreid_row_topk": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _p, _p, _p]),
    "mpreid_row_max": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "mpreid_row_kth": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _p]),
    "mpreid_row_kth_bound": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _p]),
    "mpreid_rerank_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "mpreid_rerank": (_i32, [_p, _i64, _p, _i64, _i64, _i32, _i32, _f32, _p, _i64, _p, _sz, _p, _p]),
    "mpreid_rerank_neighbor_count": (_i32, [_i32, _i32]),
    "mpreid_rerank_v0_capacity": (_i32, [_i32, _i64]),
    "mpreid_rerank_build_v0": (_i32, [_p, _i64, _p, _i64, _i64, _i32, _p, _i32, _p, _p, _p, _p, _p]),
    "mpreid_rerank_finish_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "mpreid_rerank_finish": (_i32, [_p, _i32, _p, _p, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _i32, _i32, _f32, _p, _i64, _p, _sz, _p]),
    "mpreid_dist_symmetric_topk": (_i32, [_p, _p, _p, _p, _i64, _i64, _i64, _i32, _p, _p, _p, _i64, _i64, _p, _i64, _p, _i32, _i32, _p]),
    "mpreid_cand_topk": (_i32, [_p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "mpreid_merge_topk": (_i32, [_p, _i32, _i64, _i32, _p, _p, _p, _p, _p, _p]),
    "mpreid_rerank_build_v0_sparse": (_i32, [_p, _i64, _i64, _i32, _p, _p, _i32, _p, _p, _i64, _i64, _p, _p, _p, _p, _p]),
    "mpreid_rerank_finish_ex": (_i32, [_p, _i32, _p, _p, _p, _p, _i64, _i64, _p, _p, _i64, _i64, _i64, _i32, _i32, _f32, _p, _i64, _p, _sz, _i32, _i64, _i32,
                                       _i64, _i64, _p]),
    "mpreid_rerank_finish_layout": (_i32, [_i64, _i64, _i32, _i32, C.POINTER(_i64)]),
    "mpreid_rerank_blend_default": (_i32, [_p, _i64, _i64, _p, _p, _i64, _i64, _f32, _p, _i64, _i32, _p]),
    "mpreid_hard_example_mining": (_i32, [_p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "mpreid_triplet_forward": (_i32, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "mpreid_triplet_backward": (_i32, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "mpreid_supcon_workspace_bytes": (_sz, [_i64, _i64]),
    "mpreid_supcon_step": (_i32, [_p, _i64, _i64, _i64, _p, _p, _f32, _i32, _f32, _f32, _p, _i64, _p, _i64, _i64, _p, _p, _i64, _p, _i64, _p, _sz, _p]),
    "mpreid_comm_unique_id": (_i32, [_p]),
    "mpreid_comm_init": (_i32, [C.POINTER(_p), _i32, _i32, _p]),
    "mpreid_comm_from_nccl": (_i32, [C.POINTER(_p), _p, _i32, _i32]),
    "mpreid_comm_size": (_i32, [_p, C.POINTER(_i32), C.POINTER(_i32)]),
    "mpreid_comm_broadcast": (_i32, [_p, _p, _sz, _i32, _p]),
    "mpreid_comm_allgather": (_i32, [_p, _p, _p, _sz, _p]),
    "mpreid_comm_allreduce_max_f32": (_i32, [_p, _p, _sz, _p]),
    "mpreid_comm_destroy": (_i32, [_p]),
    "mpreid_host_average_precision": (C.c_double, [_p, _i32, _i64]),
    "mpreid_host_order_keys": (None, [_p, _i64, _p]),
}


def declared_symbols() -> list[str]:
    """Every entry point include/mpreid_b200.h declares (used by the ABI test)."""
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MPREID_API[^;(]*?\b(mpreid_\w+)\s*\(", text)))


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run `python -m mp_reid_b200.build`")
        _build.build()
    elif not _build.is_current():
        # the .so is git-ignored: after an edit of csrc/ a stale build would be loaded silently (the ABI version does not
        # change with a kernel), so the source hash written at build time is checked on every load
        if build_if_missing and _build.have_nvcc():
            _build.build()
        else:
            raise RuntimeError(f"{LIB_PATH} was built from other sources than the ones in csrc/; run `python -m mp_reid_b200.build`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mpreid_abi_version() != 1:
        raise RuntimeError("libmpreid_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


class MpreidError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mpreid_last_error().decode(errors="replace")
        raise MpreidError(f"{what or 'mpreid call'} failed (code {rc}): {msg}")
