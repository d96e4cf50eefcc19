"""deflate_rs_b200 -- host-side mirror of the `deflate` crate's encode API over the B200 kernels.

Names, argument meaning and error behaviour follow the reference crate (image-rs/deflate-rs 1.0.0):

    deflate_bytes / deflate_bytes_conf            src/lib.rs:163 / :137
    deflate_bytes_zlib / deflate_bytes_zlib_conf  src/lib.rs:216 / :182
    deflate_bytes_gzip / deflate_bytes_gzip_conf  src/lib.rs:284 / :242 (GzBuilder: gzip-header 1.0)
    Compression, CompressionOptions, MatchingType, SpecialOptions
                                                  src/compression_options.rs:31-196, src/lz77.rs:26-37
    write.DeflateEncoder / write.ZlibEncoder / write.GzEncoder
                                                  src/writer.rs:89-152 / :183-290 / :331-467

Everything below this module is the C ABI of include/deflate_b200.h; all compression arithmetic runs
in sm_100a CUDA kernels (deflate-rs_b200/csrc).  There is no CPU path: importing works anywhere, but
any call that compresses raises unless the in-tree extension is built and a CUDA device is present.
"""
import ctypes
import enum
from dataclasses import dataclass

from . import _native
from ._native import DeflateB200Error, RAW, ZLIB, GZIP  # noqa: F401

__all__ = [
    "Compression", "CompressionOptions", "MatchingType", "SpecialOptions", "deflate_bytes", "deflate_bytes_conf",
    "deflate_bytes_zlib", "deflate_bytes_zlib_conf", "deflate_bytes_gzip", "deflate_bytes_gzip_conf", "GzBuilder",
    "write", "compress_device", "compress_device_batch", "compress_batch", "DeflateB200Error",
]


class MatchingType(enum.IntEnum):
    """src/lz77.rs:26-37"""
    Greedy = 0
    Lazy = 1


class SpecialOptions(enum.IntEnum):
    """src/compression_options.rs:52-59 (only Normal is implemented, as in the reference)."""
    Normal = 0


class Compression(enum.Enum):
    """src/compression_options.rs:31-42"""
    Fast = "fast"
    Default = "default"
    Best = "best"


@dataclass(frozen=True)
class CompressionOptions:
    """src/compression_options.rs:78-120; presets :126-178."""
    max_hash_checks: int = 128
    lazy_if_less_than: int = 32
    matching_type: MatchingType = MatchingType.Lazy
    special: SpecialOptions = SpecialOptions.Normal

    @staticmethod
    def default():
        return CompressionOptions(128, 32, MatchingType.Lazy)

    @staticmethod
    def high():
        return CompressionOptions(1768, 128, MatchingType.Lazy)

    @staticmethod
    def fast():
        return CompressionOptions(1, 0, MatchingType.Greedy)

    @staticmethod
    def huffman_only():
        return CompressionOptions(0, 0, MatchingType.Greedy)

    @staticmethod
    def rle():
        return CompressionOptions(0, 0, MatchingType.Lazy)

    @staticmethod
    def from_(value):
        """`impl From<Compression> for CompressionOptions` (src/compression_options.rs:188-196)."""
        if isinstance(value, CompressionOptions):
            return value
        if value is Compression.Fast:
            return CompressionOptions.fast()
        if value is Compression.Default:
            return CompressionOptions.default()
        if value is Compression.Best:
            return CompressionOptions.high()
        raise TypeError(f"expected Compression or CompressionOptions, got {value!r}")

    def _c(self):
        if not (0 <= self.max_hash_checks <= 0xFFFF and 0 <= self.lazy_if_less_than <= 0xFFFF):
            raise ValueError("max_hash_checks and lazy_if_less_than are u16 in the reference")
        return _native.dfl_options(self.max_hash_checks, self.lazy_if_less_than, int(self.matching_type), int(self.special))


class GzBuilder:
    """Host-side stand-in for `gzip_header::GzBuilder` (crate gzip-header 1.0, the type the reference's
    gzip entry points take: src/lib.rs:242, src/writer.rs:346): builds the RFC 1952 member header.
    Only the header *bytes* cross the C ABI.  Defaults: no optional fields, MTIME 0, XFL 0, OS 255."""

    def __init__(self):
        self._extra = None
        self._filename = None
        self._comment = None
        self._mtime = 0
        self._os = 255

    def mtime(self, mtime):
        self._mtime = int(mtime) & 0xFFFFFFFF
        return self

    def os(self, os_code):
        self._os = int(os_code) & 0xFF
        return self

    def extra(self, extra):
        self._extra = bytes(extra)
        return self

    def filename(self, filename):
        self._filename = bytes(filename)
        return self

    def comment(self, comment):
        self._comment = bytes(comment)
        return self

    def into_header(self) -> bytes:
        flg = (4 if self._extra is not None else 0) | (8 if self._filename is not None else 0) | \
              (16 if self._comment is not None else 0)
        h = bytearray([0x1F, 0x8B, 8, flg]) + self._mtime.to_bytes(4, "little") + bytes([0, self._os])
        if self._extra is not None:
            h += len(self._extra).to_bytes(2, "little") + self._extra
        if self._filename is not None:
            h += self._filename + b"\0"
        if self._comment is not None:
            h += self._comment + b"\0"
        return bytes(h)


def _oneshot(data, options, wrap, gz_hdr=None):
    data = bytes(data)
    opts = CompressionOptions.from_(options)._c()
    L = _native.lib()
    import numpy as np

    hdr = bytes(gz_hdr) if gz_hdr else None
    cap = L.dfl_bound(len(data), wrap) + (len(hdr) if hdr else 0)
    out = np.empty(cap, dtype=np.uint8)          # not zero-filled: the library writes every byte it reports
    n = ctypes.c_size_t()
    _native.check(L.dfl_compress(data, len(data), ctypes.byref(opts), wrap, hdr, len(hdr) if hdr else 0,
                                 out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n)), "dfl_compress")
    return out[: n.value].tobytes()


def deflate_bytes_conf(input, options):
    """src/lib.rs:137 -- raw DEFLATE stream of `input`."""
    return _oneshot(input, options, RAW)


def deflate_bytes(input):
    """src/lib.rs:163 -- Compression::Default."""
    return _oneshot(input, Compression.Default, RAW)


def deflate_bytes_zlib_conf(input, options):
    """src/lib.rs:182 -- zlib header 78 9C, stream, Adler-32 (big endian)."""
    return _oneshot(input, options, ZLIB)


def deflate_bytes_zlib(input):
    """src/lib.rs:216"""
    return _oneshot(input, Compression.Default, ZLIB)


def deflate_bytes_gzip_conf(input, options, gzip_header=None):
    """src/lib.rs:242 -- gzip member: header from `gzip_header` (a GzBuilder), stream, CRC-32 and ISIZE."""
    hdr = (gzip_header or GzBuilder()).into_header()
    return _oneshot(input, options, GZIP, hdr)


def deflate_bytes_gzip(input):
    """src/lib.rs:284"""
    return deflate_bytes_gzip_conf(input, Compression.Default, GzBuilder())


def trim() -> None:
    """dfl_trim: gives back the device scratch the library keeps between calls for the calling thread (one-shot
    context, batch pools) and the process-wide pool of parked handle resources."""
    _native.lib().dfl_trim()


def compress_device(src, options=Compression.Default, wrap=RAW, out=None, stream=None):
    """Device-resident encode: `src` and `out` are CUDA uint8 torch tensors.  Returns (out, n_bytes).

    This is dfl_compress_device: no host copies; the size is read back after the stream is synchronised.
    """
    import torch

    assert src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()
    L = _native.lib()
    n = src.numel()
    if out is None:
        out = torch.empty(L.dfl_bound(n, wrap) + 64, dtype=torch.uint8, device=src.device)
    opts = CompressionOptions.from_(options)._c()
    sz = ctypes.c_size_t()
    st = ctypes.c_void_p(stream if stream is not None else torch.cuda.current_stream(src.device).cuda_stream)
    with torch.cuda.device(src.device):
        rc = L.dfl_compress_device(ctypes.c_void_p(src.data_ptr()), n, ctypes.byref(opts), wrap, None, 0,
                                   ctypes.c_void_p(out.data_ptr()), out.numel(), ctypes.byref(sz), st)
    _native.check(rc, "dfl_compress_device")
    return out, sz.value


def compress_device_batch(srcs, options=Compression.Default, wrap=ZLIB, outs=None):
    """dfl_compress_device_batch: independent streams (a list of CUDA uint8 tensors) encoded concurrently.
    Returns (outs, sizes)."""
    import torch

    L = _native.lib()
    k = len(srcs)
    if outs is None:
        outs = [torch.empty(L.dfl_bound(s.numel(), wrap) + 64, dtype=torch.uint8, device=s.device) for s in srcs]
    opts = CompressionOptions.from_(options)._c()
    d_in = (ctypes.c_void_p * k)(*[s.data_ptr() for s in srcs])
    d_out = (ctypes.c_void_p * k)(*[o.data_ptr() for o in outs])
    n = (ctypes.c_size_t * k)(*[s.numel() for s in srcs])
    cap = (ctypes.c_size_t * k)(*[o.numel() for o in outs])
    out_len = (ctypes.c_size_t * k)()
    status = (ctypes.c_int * k)()
    torch.cuda.current_stream(srcs[0].device).synchronize() if k else None
    with torch.cuda.device(srcs[0].device if k else 0):
        rc = L.dfl_compress_device_batch(k, d_in, n, ctypes.byref(opts), wrap, d_out, cap, out_len, status)
    _native.check(rc, "dfl_compress_device_batch")
    return outs, [int(x) for x in out_len]


def compress_batch(inputs, options=Compression.Default, wrap=ZLIB):
    """dfl_compress_batch: independent streams from host memory (bytes-like objects), encoded concurrently.
    Returns a list of bytes, member i equal to deflate_bytes*_conf(inputs[i]) (lib.rs:137,182,242)."""
    import numpy as np

    L = _native.lib()
    k = len(inputs)
    if k == 0:
        return []
    srcs = [np.frombuffer(b, dtype=np.uint8) for b in inputs]
    caps = [L.dfl_bound(s.size, wrap) + 64 for s in srcs]
    outs = [np.empty(c, dtype=np.uint8) for c in caps]
    opts = CompressionOptions.from_(options)._c()
    p_in = (ctypes.c_void_p * k)(*[s.ctypes.data if s.size else None for s in srcs])
    p_out = (ctypes.c_void_p * k)(*[o.ctypes.data for o in outs])
    n = (ctypes.c_size_t * k)(*[s.size for s in srcs])
    cap = (ctypes.c_size_t * k)(*caps)
    out_len = (ctypes.c_size_t * k)()
    status = (ctypes.c_int * k)()
    rc = L.dfl_compress_batch(k, p_in, n, ctypes.byref(opts), wrap, p_out, cap, out_len, status)
    _native.check(rc, "dfl_compress_batch")
    return [o[:int(m)].tobytes() for o, m in zip(outs, out_len)]


class _Encoder:
    """Shared body of write.DeflateEncoder / write.ZlibEncoder (src/writer.rs)."""
    _wrap = RAW

    def __init__(self, writer, options, _gz_hdr=None):
        self._opts = CompressionOptions.from_(options)
        c = self._opts._c()
        self._h = _native.lib().dfl_encoder_new(ctypes.byref(c), self._wrap, _gz_hdr, len(_gz_hdr) if _gz_hdr else 0)
        if not self._h:
            raise DeflateB200Error(-1, "dfl_encoder_new")
        self._inner = writer

    # -- io::Write ---------------------------------------------------------------------------
    def write(self, buf) -> int:
        """Write::write (src/writer.rs:124-127, 254-267): returns the number of bytes consumed."""
        self._require_open()
        buf = bytes(buf)
        consumed = ctypes.c_size_t()
        _native.check(_native.lib().dfl_encoder_write(self._h, buf, len(buf), ctypes.byref(consumed)), "dfl_encoder_write")
        self._drain()          # whatever a piece produced goes to the sink now, as the reference's writer does
        return consumed.value

    def write_all(self, buf):
        buf = bytes(buf)
        while buf:
            n = self.write(buf)
            buf = buf[n:]

    def flush(self):
        """Write::flush == Z_SYNC_FLUSH (src/writer.rs:134-136): the sink then ends with 00 00 FF FF."""
        self._require_open()
        _native.check(_native.lib().dfl_encoder_flush(self._h, _native.FLUSH_SYNC), "dfl_encoder_flush")
        self._drain()

    # -- encoder specific ----------------------------------------------------------------------
    def set_piece_bytes(self, n):
        """dfl_encoder_set_piece_bytes: how much buffered input triggers an encode without a flush (the output
        does not depend on it)."""
        _native.check(_native.lib().dfl_encoder_set_piece_bytes(self._h, n), "dfl_encoder_set_piece_bytes")

    def finish(self):
        """finish(self) -> io::Result<W> (src/writer.rs:103-108, 209-214): returns the wrapped writer."""
        self._require_open()
        _native.check(_native.lib().dfl_encoder_flush(self._h, _native.FLUSH_FINISH), "dfl_encoder_flush")
        self._drain()
        w, self._inner = self._inner, None
        self._close()
        return w

    def reset(self, writer, _gz_hdr=None):
        """reset(&mut self, W) -> io::Result<W> (src/writer.rs:112-115, 218-223)."""
        self._require_open()
        _native.check(_native.lib().dfl_encoder_reset(self._h, _gz_hdr, len(_gz_hdr) if _gz_hdr else 0), "dfl_encoder_reset")
        self._drain()
        old, self._inner = self._inner, writer
        return old

    def _drain(self):
        L = _native.lib()
        p = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t()
        while True:
            _native.check(L.dfl_encoder_take_output(self._h, ctypes.byref(p), ctypes.byref(n)), "dfl_encoder_take_output")
            if n.value == 0:
                return
            chunk = ctypes.string_at(p, n.value)
            if isinstance(self._inner, bytearray):                 # the analogue of `impl Write for Vec<u8>`
                self._inner += chunk
                wrote = len(chunk)
            else:
                wrote = self._inner.write(chunk)
                wrote = len(chunk) if wrote is None else int(wrote)   # partial writes are honoured
            if wrote <= 0:
                raise IOError("failed to write whole buffer")       # io::ErrorKind::WriteZero
            L.dfl_encoder_advance_output(self._h, wrote)

    def _require_open(self):
        if not self._h:
            raise ValueError("encoder already finished")

    def _close(self):
        if self._h:
            _native.lib().dfl_encoder_free(self._h)
            self._h = None

    def __del__(self):
        # Drop (src/writer.rs:139-152): finish the stream, ignoring errors.
        try:
            if getattr(self, "_h", None) and self._inner is not None:
                self.finish()
        except Exception:
            pass
        finally:
            try:
                self._close()
            except Exception:
                pass


class DeflateEncoder(_Encoder):
    """write::DeflateEncoder<W> (src/writer.rs:89-152)."""
    _wrap = RAW


class ZlibEncoder(_Encoder):
    """write::ZlibEncoder<W> (src/writer.rs:183-290)."""
    _wrap = ZLIB

    def checksum(self) -> int:
        """Adler-32 of the data consumed so far (src/writer.rs:248)."""
        self._require_open()
        return int(_native.lib().dfl_encoder_checksum(self._h))


class GzEncoder(_Encoder):
    """write::GzEncoder<W> (src/writer.rs:331-467, feature `gzip`)."""
    _wrap = GZIP

    def __init__(self, writer, options, _gz_hdr=None):
        super().__init__(writer, options, _gz_hdr)

    @classmethod
    def from_builder(cls, builder, writer, options):
        """src/writer.rs:346-357"""
        return cls(writer, options, builder.into_header())

    def reset_with_builder(self, writer, builder):
        """src/writer.rs:403-406"""
        return self.reset(writer, builder.into_header())

    def checksum(self) -> int:
        """CRC-32 of the data consumed so far (src/writer.rs:429)."""
        self._require_open()
        return int(_native.lib().dfl_encoder_checksum(self._h))


class _WriteNamespace:
    """`deflate::write` (src/lib.rs:104-108)."""
    DeflateEncoder = DeflateEncoder
    ZlibEncoder = ZlibEncoder
    GzEncoder = GzEncoder


write = _WriteNamespace
