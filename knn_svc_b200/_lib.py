"""ctypes binding of libknnsvc_b200.so (the C ABI in include/knnsvc_b200.h).

There is no CPU fallback: if the library is missing or fails to load, importing
an op raises.  Build it with `python -m knn_svc_b200.build`.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# (KNNSVC_LIB_PATH: another build of the same library, for A/B measurements of two versions on one box)
LIB_PATH = Path(os.environ["KNNSVC_LIB_PATH"]) if os.environ.get("KNNSVC_LIB_PATH") else _PKG / "libknnsvc_b200.so"

_lib = None

i64, i32, f32, f64, vp, sz = C.c_int64, C.c_int, C.c_float, C.c_double, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/knnsvc_b200.h declares
SIGNATURES = {
    "knnsvc_last_error": (C.c_char_p, []),
    "knnsvc_version": (i32, []),
    "knnsvc_prepare_rows": (i32, [vp, i64, i32, i64, vp, i32, vp, vp, vp, vp]),
    "knnsvc_cosine_dist": (i32, [vp, i64, vp, i64, i32, vp, vp]),
    "knnsvc_knn_workspace_bytes": (sz, [i64, i64, i32, i32]),
    "knnsvc_knn_plan": (i32, [i64, i64, i32, vp]),
    "knnsvc_knn_search": (i32, [vp, vp, vp, i64, vp, vp, vp, i64, i32, i32, i32, i64, vp, vp, vp, vp, vp, sz, vp,
                                vp]),
    "knnsvc_knn_search_masked": (i32, [vp, vp, vp, i64, vp, vp, vp, i64, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp,
                                       vp, sz, vp, vp]),
    "knnsvc_knn_search_full": (i32, [vp, vp, vp, i64, vp, vp, vp, i64, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp,
                                     vp, sz, vp, vp]),
    "knnsvc_knn_workspace_layout": (i32, [i64, i64, i32, i32, vp]),
    "knnsvc_launch_count": (C.c_longlong, []),
    "knnsvc_set_option": (i32, [C.c_char_p, i32]),
    "knnsvc_filter_timing": (i32, [i32]),
    "knnsvc_filter_timing_collect": (i32, [vp, i32]),
    "knnsvc_knn_exact_workspace_bytes": (sz, [i64, i64, i32]),
    "knnsvc_knn_exact": (i32, [vp, vp, i64, vp, vp, i64, i32, i32, i64, vp, vp, vp, sz, vp]),
    "knnsvc_merge_topk": (i32, [vp, vp, i32, i64, i32, vp, vp, vp]),
    "knnsvc_merge_topk64": (i32, [vp, vp, i32, i64, i32, vp, vp, vp, vp]),
    "knnsvc_ipc_export": (i32, [vp, vp, vp]),
    "knnsvc_ipc_open": (i32, [vp, vp]),
    "knnsvc_ipc_close": (i32, [vp]),
    "knnsvc_gather_mix_sharded": (i32, [vp, vp, i32, i32, vp, vp, i64, i32, vp, vp]),
    "knnsvc_concat_cost_reselect_sharded": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, f32, vp, i32, vp, vp]),
    "knnsvc_weight_fit_sharded": (i32, [vp, vp, vp, i32, i32, vp, i32, i32, f64, i32, vp, vp, vp, sz, vp]),
    "knnsvc_gather_mix": (i32, [vp, i64, i32, vp, vp, i64, i32, vp, vp]),
    "knnsvc_f0_rerank": (i32, [vp, vp, vp, i64, i32, vp, vp]),
    "knnsvc_concat_cost_reselect": (i32, [vp, vp, vp, i64, i32, vp, vp, f32, vp, i32, vp, vp]),
    "knnsvc_weight_fit_workspace_bytes": (sz, [i64, i32]),
    "knnsvc_weight_fit": (i32, [vp, vp, i64, i32, i64, i32, f64, i32, vp, vp, vp, sz, vp]),
    "knnsvc_weight_fit_batched_workspace_bytes": (sz, [i64, i32, i32]),
    "knnsvc_weight_fit_batched": (i32, [vp, vp, i64, i32, vp, i32, i32, f64, i32, vp, vp, vp, sz, vp]),
    "knnsvc_weight_fit_amp": (i32, [vp, vp, i64, i32, vp, i32, i32, f64, i32, vp, vp, vp, vp, sz, vp]),
    "knnsvc_harmonic_bank": (i32, [vp, vp, i32, i64, i32, i32, i32, vp, vp, vp]),
    "knnsvc_layer_mix": (i32, [vp, i32, i64, i32, vp, vp, vp, vp, vp]),
    "knnsvc_stft_magnitude": (i32, [vp, i64, i64, i32, i32, vp, vp]),
    "knnsvc_harmonic_amplitudes": (i32, [vp, vp, i64, i32, i32, i32, vp, vp]),
    "knnsvc_row_l1": (i32, [vp, i64, i32, vp, vp]),
    "knnsvc_amp_ratio": (i32, [vp, vp, vp, i64, i32, i64, vp, vp]),
    "knnsvc_store_to_host": (i32, [vp, vp, C.c_size_t, vp]),
}


def load():
    """Load the shared library once and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m knn_svc_b200.build`). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    # diagnostic switches from the environment: KNNSVC_OPTIONS="concat_cluster=0,refine_min_candidates=800"
    for item in filter(None, os.environ.get("KNNSVC_OPTIONS", "").split(",")):
        name, _, value = item.partition("=")
        if lib.knnsvc_set_option(name.strip().encode(), int(value)) != 0:
            raise ValueError(f"KNNSVC_OPTIONS: {lib.knnsvc_last_error().decode(errors='replace')}")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().knnsvc_last_error().decode(errors="replace")
        if rc < 0:
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")
