"""ctypes binding of libpraline_b200.so (C ABI: include/praline_b200.h).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded this module
raises, and every entry point that needs a device fails loudly when there is none.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGPU_LIB") or os.path.join(_HERE, "libpraline_b200.so")   # PGPU_LIB: experiment builds

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

_SIGNATURES = {
    "pgpu_abi_version": (c_int, []),
    "pgpu_init": (c_int, [c_int]),
    "pgpu_shutdown": (None, []),
    "pgpu_last_error": (ctypes.c_char_p, []),
    "pgpu_supported_k": (c_int, [c_int]),
    "pgpu_warps_per_tile": (c_int, []),
    "pgpu_align_batch": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                                 c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgpu_align_profiles": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgpu_align_tiles": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                 c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_float, c_float, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "pgpu_align_tiles16": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                   c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pgpu_traceback_tiles": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                     c_float, c_int, c_void_p]),
    "pgpu_align_tiles16_traced": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                          c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p]),
    "pgpu_align_tiles16_paired_traced": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                                 c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_void_p, c_void_p]),
    "pgpu_traceback_dual": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "pgpu_align_tiles_local": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int,
                                       c_float, c_float, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgpu_traceback_tiles_local": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int,
                                           c_void_p]),
    "pgpu_build_scores": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                  c_void_p]),
    "pgpu_build_rows": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p]),
    "pgpu_profile_times_matrix": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "pgpu_build_rows_fast": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "pgpu_build_rows_tc": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "pgpu_split_residents": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pgpu_build_scores_seq": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                      c_void_p]),
    "pgpu_general_workspace_bytes": (c_int64, [c_int, c_int]),
    "pgpu_align_general": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "pgpu_align_profile_long": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                        c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgpu_counts_to_profile": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "pgpu_merge_counts": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "pgpu_fill_debug": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "pgpu_microbench": (c_int, [c_void_p, c_int]),
    "pgpu_cluster_workspace_bytes": (c_int64, [c_int]),
    "pgpu_cluster_merge_order": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgpu_tree_distance": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgpu_plan_profile_wave": (ctypes.c_longlong, [c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_int, c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_int,
                                                   c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)


class PralineGpuError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (no device is touched)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PralineGpuError(
                "CUDA library %s not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.pgpu_abi_version() != 1:
            raise PralineGpuError("ABI version mismatch in %s" % LIB_PATH)
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise PralineGpuError("libpraline_b200: %s (code %d)" % (load().pgpu_last_error().decode(), rc))
