"""ctypes binding of tests/model/libdfl_model.so -- TEST INFRASTRUCTURE ONLY (CPU walk-through of
the GPU pipeline's algorithms built from deflate-rs_b200/csrc/dfl_core.h)."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model")
_SO = os.path.join(_DIR, "libdfl_model.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR, "-s"])
        L = ctypes.CDLL(_SO)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.dflm_compress.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint16, ctypes.c_uint16, ctypes.c_uint8,
                                    ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(u8p),
                                    ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_uint32)]
        L.dflm_tokens.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint16, ctypes.c_uint16, ctypes.c_uint8,
                                  ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                  ctypes.POINTER(ctypes.POINTER(ctypes.c_uint32)), ctypes.POINTER(ctypes.c_size_t)]
        L.dflm_symbols.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
        L.dflm_free.argtypes = [ctypes.c_void_p]
        L.dflm_crc32_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
        L.dflm_crc32_combine.restype = ctypes.c_uint32
        L.dflm_adler32_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
        L.dflm_adler32_combine.restype = ctypes.c_uint32
        L.dflm_resolve_stats.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
        _lib = L
    return _lib


def force_generic_match(on: bool):
    """Test hook: run the generic candidate walk also for max_hash_checks == 1 (the kernels and the model take the
    one-candidate path there: the match is settled against the predecessor in the sorted order)."""
    lib().dflm_force_generic_match(1 if on else 0)


def resolve_stats(reset=True):
    """Counters of the long-match resolutions the parser asked for since the last reset."""
    st = (ctypes.c_uint64 * 8)()
    lib().dflm_resolve_stats(st, 1 if reset else 0)
    return {"resolutions": st[0], "candidates": st[1], "bytes": st[2], "visits": st[3], "nearest_is_answer": st[4],
            "floor_ge_8": st[5], "no_result": st[6]}


def compress(data, opts, pseg=8192, warm=1024, rounds=4):
    out = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_size_t()
    stats = (ctypes.c_uint32 * 4)()
    lib().dflm_compress(data, len(data), opts.max_hash_checks, opts.lazy_if_less_than, opts.matching_type, pseg, warm,
                        rounds, ctypes.byref(out), ctypes.byref(n), stats)
    res = ctypes.string_at(out, n.value)
    lib().dflm_free(out)
    return res, {"repairs": stats[0], "seq_repairs": stats[1], "tokens": stats[2]}


def tokens(data, opts, pseg=8192, warm=1024, rounds=4):
    out = ctypes.POINTER(ctypes.c_uint32)()
    n = ctypes.c_size_t()
    lib().dflm_tokens(data, len(data), opts.max_hash_checks, opts.lazy_if_less_than, opts.matching_type, pseg, warm,
                      rounds, ctypes.byref(out), ctypes.byref(n))
    arr = np.ctypeslib.as_array(out, shape=(max(n.value, 1),))[: n.value].copy()
    lib().dflm_free(out)
    return arr


def symbols(length, dist):
    o = (ctypes.c_uint32 * 6)()
    lib().dflm_symbols(length, dist, o)
    return list(o)
