"""ctypes binding of ``include/radad_flat.h`` (the C ABI of ``libradad_flat.so``).

This is the same stub a maintainer of the reference would add next to ``vector_database.py`` in
place of ``import faiss``.  There is deliberately no fallback: if the shared
library is missing or a symbol is absent, importing the product fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libradad_flat.so")

METRIC_L2, METRIC_IP = 0, 1
STORE_F32, STORE_BF16, STORE_F16 = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
ALGO_AUTO, ALGO_SIMT, ALGO_TC, ALGO_STREAM = 0, 1, 2, 3
FLAG_KEEP_F32_MASTER = 1
ABI_VERSION = 2

_fp = POINTER(c_float)
_ip = POINTER(c_int64)
_h = c_void_p

# name -> (restype, argtypes); must list EVERY symbol include/radad_flat.h declares (tests check this).
SIGNATURES = {
    "rdb_abi_version": (c_int, []),
    "rdb_create": (c_int, [c_int, c_int, c_int, c_int, c_uint, POINTER(_h)]),
    "rdb_destroy": (c_int, [_h]),
    "rdb_last_error": (c_char_p, [_h]),
    "rdb_set_stream": (c_int, [_h, c_void_p]),
    "rdb_use_own_stream": (c_int, [_h]),
    "rdb_sync": (c_int, [_h]),
    "rdb_reserve": (c_int, [_h, c_int64]),
    "rdb_add": (c_int, [_h, c_void_p, c_int64, c_int, c_int]),
    "rdb_search": (c_int, [_h, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rdb_search_algo": (c_int, [_h, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rdb_search_shard": (c_int, [_h, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdb_merge_shards": (c_int, [_h, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "rdb_ipc_alloc": (c_int, [_h, c_size_t, POINTER(c_void_p), c_void_p]),
    "rdb_ipc_open": (c_int, [_h, c_void_p, POINTER(c_void_p)]),
    "rdb_ipc_close": (c_int, [_h, c_void_p]),
    "rdb_ipc_free": (c_int, [_h, c_void_p]),
    "rdb_merge_shards_peer": (c_int, [_h, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_int, c_int64,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdb_enable_peer_access": (c_int, [_h, c_int]),
    "rdb_copy_async": (c_int, [_h, c_void_p, c_void_p, c_size_t]),
    "rdb_reconstruct": (c_int, [_h, c_int64, c_void_p]),
    "rdb_reconstruct_batch": (c_int, [_h, c_void_p, c_int64, c_int, c_void_p]),
    "rdb_filter_first_k": (c_int, [_h, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p]),
    "rdb_set_labels": (c_int, [_h, c_void_p, c_int64]),
    "rdb_label_vote": (c_int, [_h, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "rdb_ntotal": (c_int64, [_h]),
    "rdb_dim": (c_int, [_h]),
    "rdb_metric": (c_int, [_h]),
    "rdb_store_dtype": (c_int, [_h]),
    "rdb_set_id_offset": (c_int, [_h, c_int64]),
    "rdb_serialize": (c_int, [_h, c_char_p]),
    "rdb_deserialize": (c_int, [c_char_p, c_int, c_int, c_uint, POINTER(_h)]),
    "rdb_mem_info": (c_int, [_h, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)]),
    "rdb_release_scratch": (c_int, [_h]),
    "rdb_truncate": (c_int, [_h, c_int64]),
    "rdb_set_option": (c_int, [_h, c_char_p, c_int64]),
    "rdb_launch_count": (c_int64, [_h]),
    "rdb_host_sync_count": (c_int64, [_h]),
    "rdb_last_kernel_ms": (c_int, [_h, POINTER(c_float), POINTER(c_int), POINTER(c_int)]),
    "rdb_last_uncertified": (c_int64, [_h]),
    "rdb_last_tier1": (c_int, [_h, POINTER(c_int64), POINTER(c_int64), POINTER(c_int)]),
}

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def load() -> ctypes.CDLL:
    """Load ``libradad_flat.so`` (built in-tree by ``csrc/build.sh`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with csrc/build.sh (nvcc, sm_100a). "
            "This package has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.rdb_abi_version() != ABI_VERSION:
        raise NativeLibraryMissing(f"{LIB_PATH}: ABI version {lib.rdb_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def last_error(handle) -> str:
    msg = load().rdb_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, handle=None) -> None:
    """Non-zero status -> RuntimeError (what SWIG-faiss raises for a failed C++ call)."""
    if rc != 0:
        raise RuntimeError(f"radad_flat error {rc}: {last_error(handle)}")
