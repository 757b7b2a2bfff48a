"""ctypes binding of libanorag_b200.so (C ABI declared in include/anorag_b200.h).

There is no CPU fallback: if the shared object is missing (and cannot be built
because nvcc is absent) importing this module raises, and every compute call on a
machine without a CUDA device fails with the library's ARB_ECUDA error.
ctypes.CDLL releases the GIL for the duration of each call, so concurrent
searches from Python thread pools (reference main_musique.py:487,
query/query_processor.py:2761) overlap on the device.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libanorag_b200.so")

ARB_F32, ARB_BF16, ARB_F64 = 0, 1, 2
STORE_BF16, STORE_FP32_SPLIT = 0, 1
FUSE_LINEAR, FUSE_RRF = 0, 1
MAX_TOPK = 256
MAX_QUERY_TERMS = 64


class NativeError(RuntimeError):
    """A libanorag_b200 entry point returned a non-zero status."""

    def __init__(self, code, message):
        super().__init__(f"libanorag_b200 error {code}: {message}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        # in-tree build on first use (nvcc cross-compiles without a GPU)
        import importlib.util
        spec = importlib.util.spec_from_file_location("_arb_build", os.path.join(_HERE, "csrc", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    try:                                # torch maps libcudart.so.12, which the library links dynamically; the image's
        import torch  # noqa: F401      # own /usr/local/cuda copy serves when torch is absent
    except Exception:
        pass
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_i64p = POINTER(c_int64)
_SIGS = {
    "arb_abi_version": (c_int, []),
    "arb_last_error": (c_char_p, []),
    "arb_device_count": (c_int, []),
    "arb_launch_count": (c_int64, [c_int]),
    "arb_profile": (c_int, [c_int]),
    "arb_profile_collect": (c_int, [POINTER(c_double), POINTER(c_int64)]),
    "arb_index_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int]),
    "arb_index_destroy": (c_int, [c_void_p]),
    "arb_index_add": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "arb_index_reset": (c_int, [c_void_p]),
    "arb_index_ntotal": (c_int64, [c_void_p]),
    "arb_index_set_row_normalization": (c_int, [c_void_p, c_int]),
    "arb_index_dim": (c_int, [c_void_p]),
    "arb_index_reconstruct": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "arb_index_search": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int64,
                                 c_void_p, c_void_p, c_int, c_void_p]),
    "arb_index_score_rows": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                     c_int, c_void_p]),
    "arb_index_ivf_set_centroids": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "arb_index_ivf_train": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p]),
    "arb_index_ivf_nlist": (c_int, [c_void_p]),
    "arb_index_ivf_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "arb_index_ivf_search": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int, c_int64, c_void_p,
                                     c_void_p, c_int, c_void_p]),
    "arb_bm25_create": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int64, c_void_p]),
    "arb_bm25_destroy": (c_int, [c_void_p]),
    "arb_bm25_ndocs": (c_int64, [c_void_p]),
    "arb_bm25_npostings": (c_int64, [c_void_p]),
    "arb_bm25_query_postings": (c_int64, [c_void_p, c_void_p, c_int64, c_int]),
    "arb_bm25_search": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p, c_void_p,
                                c_int, c_void_p]),
    "arb_bm25_scores": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p, c_int,
                                c_void_p]),
    "arb_bm25_batch_stats": (c_int, [POINTER(c_int64), c_int]),
    "arb_combine_fields": (c_int, [c_int, POINTER(c_void_p), POINTER(c_double), c_int64, c_int64, c_void_p, c_void_p]),
    "arb_fuse": (c_int, [c_int, POINTER(c_double), c_double, c_int, c_int64, POINTER(c_void_p), POINTER(c_void_p),
                         POINTER(c_int), POINTER(c_int), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "arb_rescore_candidates": (c_int, [c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_double, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "arb_rescore_v2": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_double, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "arb_rescore_khop": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_double, c_double,
                                 c_int, c_void_p, c_void_p, c_void_p]),
    "arb_tfidf_create": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "arb_tfidf_destroy": (c_int, [c_void_p]),
    "arb_tfidf_search": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "arb_merge_topk": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p,
                               c_void_p]),
    "arb_merge_topk_peers": (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), c_int, c_int64, c_int64, c_int,
                                     c_void_p, c_void_p, c_void_p]),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here = ABI mismatch, fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED = tuple(_SIGS)


def last_error() -> str:
    msg = lib.arb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    if rc != 0:
        raise NativeError(rc, last_error())


def device_count() -> int:
    n = lib.arb_device_count()
    if n < 0:
        raise NativeError(n, last_error())
    return n


def launch_count(reset: bool = False) -> int:
    return int(lib.arb_launch_count(1 if reset else 0))


PROFILE_KINDS = ("dense_topk", "bm25_topk", "fuse", "dense_merge", "bm25_sparse", "bm25_planes", "bm25_final",
                 "ivf_coarse")


def profile(enable: bool) -> None:
    check(lib.arb_profile(1 if enable else 0))


def profile_collect():
    """-> {kind: (total_ms, launches)} for the kernels bracketed since profile(True)."""
    ms = (c_double * len(PROFILE_KINDS))()
    cnt = (c_int64 * len(PROFILE_KINDS))()
    check(lib.arb_profile_collect(ms, cnt))
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(PROFILE_KINDS)}


def bm25_batch_stats(reset: bool = True):
    """Diagnostics of the batched BM25 pipeline (needs ARB_BM25_STATS=1 in the environment)."""
    out = (c_int64 * 8)()
    check(lib.arb_bm25_batch_stats(out, 1 if reset else 0))
    return {"queries": int(out[0]), "fallback_queries": int(out[1]), "emissions": int(out[2]),
            "max_emissions": int(out[3]), "fp64_rechecks": int(out[4]), "set_aside_pairs": int(out[5]),
            "set_aside_matched": int(out[6])}
