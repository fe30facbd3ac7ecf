"""ctypes binding of the C-ABI in include/kiez_b200.h (libkiez_b200.so).

There is no CPU fallback: if the shared library has not been built, importing
this module raises ImportError; if no CUDA device is present, constructing the
``B200`` backend raises ImportError (the contract kiez uses to probe backends,
kiez/kiez.py:118-122, kiez/neighbors/util.py:31-38).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KB2_LIB selects another build of the same library (kernel A/B experiments only)
LIB_PATH = os.environ.get("KB2_LIB") or os.path.join(_HERE, "lib", "libkiez_b200.so")

METRIC_EUCLIDEAN, METRIC_SQEUCLIDEAN, METRIC_COSINE = 0, 1, 2
RESCALE_CSLS, RESCALE_LS, RESCALE_NICDM, RESCALE_MP_GAUSS = 0, 1, 2, 3
KNN_AUTO, KNN_TC, KNN_SIMT, KNN_TC1 = 0, 1, 2, 3
#: select.cuh EMPTY_ENTRY (key +inf, id -1) as a signed 64-bit value: pads packed candidate entries
EMPTY_ENTRY = 0xFF800000FFFFFFFF - (1 << 64)

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double

# name -> argtypes, mirrors include/kiez_b200.h one to one
SIGNATURES = {
    "kb2_version": [],
    "kb2_max_candidates": [],
    "kb2_padded_dim": [_int],
    "kb2_suggest_splits": [_i64, _i64, _int, _int],
    "kb2_prepare_rows": [_p, _i64, _int, _i64, _p, _int, _p, _p, _int, _p, _p, _p],
    "kb2_knn_candidates": [_int, _p, _p, _i64, _p, _p, _p, _i64, _int, _int, _int, _p, _p, _p],
    "kb2_knn_fused": [_p, _p, _p, _i64, _p, _p, _p, _i64, _int, _int, _int, _p, _p, _p, _int, _i64,
                      _p, _p],
    "kb2_col_select": [_p, _p, _i64, _int, _int, _p, _p, _p, _p, _p],
    "kb2_col_compact": [_p, _p, _i64, _int, _int, _p, _p],
    "kb2_col_heads": [_p, _p, _i64, _int, _int, _i64, _p, _p, _p],
    "kb2_kth_key": [_p, _int, _i64, _i64, _int, _int, _p, _p],
    "kb2_screen_stages": [_int, _int, _int],
    "kb2_screen_config": [_int, _int, _int, _int, _i64, _p, _p],
    "kb2_screen_plan": [_i64, _i64, _int, _int, _int, _p, _p],
    "kb2_knn_screen": [_p, _p, _i64, _p, _p, _i64, _int, _int, _int, _int, _p, _p, _p, _p, _p, _p,
                       _int, _i64, _p],
    "kb2_max_f32": [_p, _i64, _p, _p],
    "kb2_split_error_terms": [_p, _i64, _int, _p, _p, _p],
    "kb2_refine_topk_checked": [_p, _i64, _i64, _p, _i64, _i64, _int, _int, _p, _p, _p, _int, _int,
                                _i64, _int, _i64, _int, _p, _p, _p, _i64, _int, _int, _p, _p, _p,
                                _p, _dbl, _p, _p],
    "kb2_refine_topk": [_p, _i64, _i64, _p, _i64, _i64, _int, _int, _p, _p, _p, _int, _int, _i64,
                        _int, _i64, _int, _p, _p, _p],
    "kb2_topk_rows": [_p, _p, _i64, _int, _int, _i64, _int, _p, _p, _p],
    "kb2_row_stats": [_p, _i64, _int, _p, _p, _p, _p],
    "kb2_rescale_topk": [_int, _p, _p, _i64, _int, _p, _p, _i64, _int, _p, _p, _p],
    "kb2_mp_empiric_topk": [_p, _p, _i64, _int, _p, _p, _i64, _int, _int, _p, _p, _p],
    "kb2_dsl_fit": [_p, _i64, _i64, _p, _i64, _i64, _int, _int, _p, _int, _p, _p, _p],
    "kb2_dsl_transform": [_p, _i64, _i64, _p, _i64, _i64, _int, _int, _p, _int, _p, _p, _p, _p],
    "kb2_dsl_finish_topk": [_p, _p, _i64, _int, _p, _int, _int, _p, _p, _p],
    "kb2_index_range": [_p, _i64, _i64, _int, _p, _p],
    "kb2_k_occurrence": [_p, _i64, _i64, _int, _i64, _p, _p],
    "kb2_hub_moments": [_p, _i64, _dbl, _dbl, _p, _p],
    "kb2_compact_ids": [_p, _i64, _int, _dbl, _p, _p, _p, _p],
    "kb2_gini_numerator": [_p, _i64, _i64, _p, _p, _p],
    "kb2_hits": [_p, _i64, _i64, _int, _p, _p, _int, _p, _p],
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA library first "
        "(python -c 'import __graft_entry__ as g; g.build()' or kiez_b200/csrc/build.sh). "
        "kiez_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)
lib.kb2_last_error.restype = C.c_char_p
lib.kb2_last_error.argtypes = []
for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _int
    _fn.argtypes = _args

#: kernels each entry point launches (memsets not counted); bench.py reports the total
KERNELS_PER_CALL = {"kb2_index_range": 2, "kb2_compact_ids": 3, "kb2_gini_numerator": 2,
                    "kb2_split_error_terms": 2}
#: number of kiez_b200 kernels launched through `call` so far
launch_counter = 0


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point; raise RuntimeError on failure."""
    global launch_counter
    launch_counter += KERNELS_PER_CALL.get(name, 1)
    status = getattr(lib, name)(*args)
    if status != 0:
        raise RuntimeError(f"{name} failed: {lib.kb2_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
