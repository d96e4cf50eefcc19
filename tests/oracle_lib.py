"""ctypes binding of oracle/libdeflate_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's encode path (oracle/deflate_oracle.c).
Nothing in the product package imports this module.
"""
import ctypes
import os
import subprocess

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "libdeflate_oracle.so")

RAW, ZLIB, GZIP = 0, 1, 2


class Options(ctypes.Structure):
    """compression_options.rs:78-120"""

    _fields_ = [
        ("max_hash_checks", ctypes.c_uint16),
        ("lazy_if_less_than", ctypes.c_uint16),
        ("matching_type", ctypes.c_uint8),
        ("special", ctypes.c_uint8),
    ]


class Token(ctypes.Structure):
    _fields_ = [("dist", ctypes.c_uint16), ("litlen", ctypes.c_uint8), ("pad", ctypes.c_uint8)]


# presets, compression_options.rs:14-20,126-178
def opts_default():
    return Options(128, 32, 1, 0)


def opts_fast():
    return Options(1, 0, 0, 0)


def opts_high():
    return Options(1768, 128, 1, 0)


def opts_rle():
    return Options(0, 0, 1, 0)


def opts_huffman_only():
    return Options(0, 0, 0, 0)


PRESETS = {
    "default": opts_default,
    "fast": opts_fast,
    "high": opts_high,
    "rle": opts_rle,
    "huffman_only": opts_huffman_only,
}


def build():
    src = os.path.join(_ORACLE_DIR, "deflate_oracle.c")
    hdr = os.path.join(_ORACLE_DIR, "deflate_oracle.h")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = ctypes.CDLL(build())
    u8p = ctypes.POINTER(ctypes.c_uint8)
    L.dfo_compress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(Options), ctypes.c_int,
                               ctypes.POINTER(u8p), ctypes.POINTER(ctypes.c_size_t)]
    L.dfo_compress.restype = ctypes.c_int
    L.dfo_free.argtypes = [ctypes.c_void_p]
    L.dfo_stream_new.argtypes = [ctypes.POINTER(Options), ctypes.c_int]
    L.dfo_stream_new.restype = ctypes.c_void_p
    for name in ("dfo_stream_flush", "dfo_stream_finish", "dfo_stream_reset"):
        getattr(L, name).argtypes = [ctypes.c_void_p]
        getattr(L, name).restype = ctypes.c_int
    L.dfo_stream_write.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    L.dfo_stream_write.restype = ctypes.c_int
    L.dfo_stream_checksum.argtypes = [ctypes.c_void_p]
    L.dfo_stream_checksum.restype = ctypes.c_uint32
    L.dfo_stream_output.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
    L.dfo_stream_output.restype = u8p
    L.dfo_stream_clear_output.argtypes = [ctypes.c_void_p]
    L.dfo_stream_free.argtypes = [ctypes.c_void_p]
    L.dfo_lz77_tokens.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(Options),
                                  ctypes.POINTER(ctypes.POINTER(Token)), ctypes.POINTER(ctypes.c_size_t),
                                  ctypes.POINTER(ctypes.POINTER(ctypes.c_size_t)), ctypes.POINTER(ctypes.c_size_t)]
    L.dfo_get_match_length.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t]
    L.dfo_get_match_length.restype = ctypes.c_size_t
    L.dfo_longest_match_filled.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                           ctypes.c_size_t, ctypes.c_uint16, ctypes.POINTER(ctypes.c_size_t),
                                           ctypes.POINTER(ctypes.c_size_t)]
    L.dfo_huffman_lengths.argtypes = [ctypes.POINTER(ctypes.c_uint16), ctypes.c_size_t, ctypes.c_size_t, u8p]
    L.dfo_encode_lengths.argtypes = [u8p, ctypes.c_size_t, u8p, u8p, ctypes.POINTER(ctypes.c_uint16)]
    L.dfo_encode_lengths.restype = ctypes.c_size_t
    L.dfo_create_codes.argtypes = [u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint16)]
    L.dfo_length_code.argtypes = [ctypes.c_uint16, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]
    L.dfo_length_code.restype = ctypes.c_uint
    L.dfo_distance_code.argtypes = [ctypes.c_uint16, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]
    L.dfo_distance_code.restype = ctypes.c_uint
    L.dfo_bitwriter_kat.argtypes = [ctypes.POINTER(ctypes.c_uint16), u8p, ctypes.c_size_t, u8p, ctypes.c_size_t]
    L.dfo_bitwriter_kat.restype = ctypes.c_size_t
    L.dfo_stored_padding.argtypes = [ctypes.c_uint8]
    L.dfo_stored_padding.restype = ctypes.c_uint64
    L.dfo_reverse_bits.argtypes = [ctypes.c_uint16, ctypes.c_uint8]
    L.dfo_reverse_bits.restype = ctypes.c_uint16
    L.dfo_compress_fixed.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(u8p),
                                     ctypes.POINTER(ctypes.c_size_t)]
    L.dfo_adler32.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
    L.dfo_adler32.restype = ctypes.c_uint32
    L.dfo_crc32.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t]
    L.dfo_crc32.restype = ctypes.c_uint32
    L.dfo_hash_table_filled.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint16),
                                        ctypes.POINTER(ctypes.c_uint16)]
    _lib = L
    return L


def _take(ptr, n):
    out = ctypes.string_at(ptr, n) if n else b""
    lib().dfo_free(ptr)
    return out


def compress(data: bytes, opts: Options = None, wrap: int = RAW) -> bytes:
    """deflate_bytes_conf / deflate_bytes_zlib_conf / deflate_bytes_gzip_conf (lib.rs:137,182,242)."""
    opts = opts or opts_default()
    out = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_size_t()
    rc = lib().dfo_compress(data, len(data), ctypes.byref(opts), wrap, ctypes.byref(out), ctypes.byref(n))
    assert rc == 0
    return _take(out, n.value)


def compress_fixed(data: bytes) -> bytes:
    out = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_size_t()
    lib().dfo_compress_fixed(data, len(data), ctypes.byref(out), ctypes.byref(n))
    return _take(out, n.value)


def lz77_tokens(data: bytes, opts: Options):
    """Returns (list of (litlen, dist), list of cumulative block ends)."""
    toks = ctypes.POINTER(Token)()
    n = ctypes.c_size_t()
    be = ctypes.POINTER(ctypes.c_size_t)()
    nb = ctypes.c_size_t()
    lib().dfo_lz77_tokens(data, len(data), ctypes.byref(opts), ctypes.byref(toks), ctypes.byref(n),
                          ctypes.byref(be), ctypes.byref(nb))
    import numpy as np

    arr = np.ctypeslib.as_array(ctypes.cast(toks, ctypes.POINTER(ctypes.c_uint32)), shape=(max(n.value, 1),))[: n.value].copy()
    ends = [be[i] for i in range(nb.value)]
    lib().dfo_free(toks)
    lib().dfo_free(be)
    dist = (arr & 0xFFFF).astype("uint32")
    litlen = ((arr >> 16) & 0xFF).astype("uint32")
    return litlen, dist, ends


class Stream:
    """write::{DeflateEncoder,ZlibEncoder,GzEncoder} over a Vec sink (writer.rs:89-290)."""

    def __init__(self, opts: Options = None, wrap: int = RAW):
        self._opts = opts or opts_default()
        self._h = lib().dfo_stream_new(ctypes.byref(self._opts), wrap)

    def write(self, data: bytes):
        assert lib().dfo_stream_write(self._h, data, len(data)) == 0

    def flush(self):
        lib().dfo_stream_flush(self._h)

    def finish(self) -> bytes:
        lib().dfo_stream_finish(self._h)
        return self.output()

    def reset(self) -> bytes:
        lib().dfo_stream_reset(self._h)
        out = self.output()
        lib().dfo_stream_clear_output(self._h)
        return out

    def checksum(self) -> int:
        return lib().dfo_stream_checksum(self._h)

    def output(self) -> bytes:
        n = ctypes.c_size_t()
        p = lib().dfo_stream_output(self._h, ctypes.byref(n))
        return ctypes.string_at(p, n.value) if n.value else b""

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dfo_stream_free(self._h)
            self._h = None
