"""ctypes binding of libdeflate_b200.so (the C ABI declared in include/deflate_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
fallback of any kind: if the library is missing, or no CUDA device is present, calls raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DFL_LIB_PATH: tuning hook (tools/tune_variants.sh builds several libraries with different kernel constants)
LIB_PATH = os.environ.get("DFL_LIB_PATH") or os.path.join(_HERE, "libdeflate_b200.so")

RAW, ZLIB, GZIP = 0, 1, 2
FLUSH_SYNC, FLUSH_FINISH = 1, 2
OK, AGAIN = 0, 1
E_OVERFLOW = -5
E_NCCL = -10
COMM_ID_BYTES = 128

# every symbol include/deflate_b200.h declares
EXPORTS = [
    "dfl_options_preset", "dfl_strerror", "dfl_last_cuda_error", "dfl_version", "dfl_device_count", "dfl_bound",
    "dfl_compress", "dfl_compress_device", "dfl_compress_device_piece", "dfl_compress_device_batch", "dfl_compress_batch", "dfl_set_profiling", "dfl_last_stage_times", "dfl_last_counters", "dfl_trim",
    "dfl_encoder_new", "dfl_encoder_write", "dfl_encoder_flush", "dfl_encoder_set_piece_bytes", "dfl_encoder_take_output",
    "dfl_encoder_advance_output", "dfl_encoder_checksum", "dfl_encoder_reset", "dfl_encoder_free",
    "dfl_adler32_device", "dfl_crc32_device", "dfl_encode_tokens", "dfl_lz77_tokens",
    "dfl_comm_unique_id", "dfl_comm_init", "dfl_comm_free", "dfl_comm_world", "dfl_comm_rank", "dfl_comm_last_error",
    "dfl_gather_device",
]


class dfl_options(ctypes.Structure):
    _fields_ = [
        ("max_hash_checks", ctypes.c_uint16),
        ("lazy_if_less_than", ctypes.c_uint16),
        ("matching_type", ctypes.c_uint8),
        ("special", ctypes.c_uint8),
    ]


class DeflateB200Error(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib().dfl_strerror(status).decode()
        detail = lib().dfl_last_cuda_error().decode()
        super().__init__(f"{where}: {msg} (status {status})" + (f" [{detail}]" if detail else ""))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). deflate_rs_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    optp = ctypes.POINTER(dfl_options)
    szp = ctypes.POINTER(ctypes.c_size_t)
    L.dfl_options_preset.argtypes = [ctypes.c_int, optp]
    L.dfl_strerror.argtypes = [ctypes.c_int]
    L.dfl_strerror.restype = ctypes.c_char_p
    L.dfl_last_cuda_error.restype = ctypes.c_char_p
    L.dfl_bound.argtypes = [ctypes.c_size_t, ctypes.c_int]
    L.dfl_bound.restype = ctypes.c_size_t
    L.dfl_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, optp, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                               ctypes.c_void_p, ctypes.c_size_t, szp]
    L.dfl_compress_device.argtypes = [ctypes.c_void_p, ctypes.c_size_t, optp, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, szp, ctypes.c_void_p]
    L.dfl_compress_device_piece.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, optp, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_size_t, szp, ctypes.c_void_p]
    L.dfl_compress_device_batch.argtypes = [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), szp, optp, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_void_p), szp, szp, ctypes.POINTER(ctypes.c_int)]
    L.dfl_compress_batch.argtypes = [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), szp, optp, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_void_p), szp, szp, ctypes.POINTER(ctypes.c_int)]
    L.dfl_set_profiling.argtypes = [ctypes.c_int]
    L.dfl_last_stage_times.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    L.dfl_last_counters.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
    L.dfl_encoder_new.argtypes = [optp, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    L.dfl_encoder_new.restype = ctypes.c_void_p
    L.dfl_encoder_write.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, szp]
    L.dfl_encoder_flush.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.dfl_encoder_set_piece_bytes.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.dfl_encoder_take_output.argtypes = [ctypes.c_void_p, ctypes.POINTER(u8p), szp]
    L.dfl_encoder_advance_output.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.dfl_encoder_advance_output.restype = None
    L.dfl_encoder_checksum.argtypes = [ctypes.c_void_p]
    L.dfl_encoder_checksum.restype = ctypes.c_uint32
    L.dfl_encoder_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.dfl_encoder_free.argtypes = [ctypes.c_void_p]
    L.dfl_encoder_free.restype = None
    L.dfl_adler32_device.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p]
    L.dfl_crc32_device.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p]
    L.dfl_encode_tokens.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                    ctypes.c_size_t, szp]
    L.dfl_lz77_tokens.argtypes = [ctypes.c_void_p, ctypes.c_size_t, optp, ctypes.c_void_p, ctypes.c_size_t, szp]
    L.dfl_comm_unique_id.argtypes = [ctypes.c_void_p]
    L.dfl_comm_init.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.dfl_comm_free.argtypes = [ctypes.c_void_p]
    L.dfl_comm_free.restype = None
    L.dfl_comm_world.argtypes = [ctypes.c_void_p]
    L.dfl_comm_rank.argtypes = [ctypes.c_void_p]
    L.dfl_comm_last_error.restype = ctypes.c_char_p
    L.dfl_gather_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, szp,
                                    ctypes.c_int, ctypes.c_void_p]
    _lib = L
    return L


def check(status, where):
    if status != OK:
        raise DeflateB200Error(status, where)
