"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle.

Bar (DESIGN.md "parity"): the GPU stream is BYTE-IDENTICAL to the oracle's (= the reference
algorithm's) for one-shot calls at every preset, and always inflates to the input with an
independent inflater (CPython zlib).  Needs a B200: run with `pytest -m gpu`."""
import ctypes
import os
import zlib

import numpy as np
import pytest

import oracle_lib as o
from conftest import FIXTURES, fixture_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dfl():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import deflate_rs_b200 as d
    assert d._native.lib().dfl_device_count() >= 1
    return d


def _copts(dfl, opts):
    return dfl.CompressionOptions(opts.max_hash_checks, opts.lazy_if_less_than, dfl.MatchingType(opts.matching_type))


def _inputs(pg11):
    rng = np.random.default_rng(1)
    d = {
        "pg11": pg11, "short": fixture_bytes("short.bin"), "issue18": fixture_bytes("issue_18_201911.bin"),
        "dump": fixture_bytes("dump.bin"), "zeros65537": bytes(65537), "zeros61000": bytes(61000),
        "lastblock": bytes([22]) * 32768 + bytes([5, 2, 55, 11, 12]), "fives": bytes([5]) * 100000,
        "empty": b"", "one": b"\x01", "four": bytes([5, 6, 7, 8]), "six": bytes([10, 10, 10, 10, 10, 55]),
        "random": rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(),
        "random4": rng.integers(0, 4, 200000, dtype=np.uint8).tobytes(),
        "period7": bytes(range(7)) * 30000,
        "gnu": b"                    GNU GENERAL PUBLIC LICENSE",
    }
    for n in (2, 3, 5, 259, 32767, 32768, 32769, 65535, 65536, 65537, 65794, 65795):
        d[f"pg{n}"] = pg11[:n]
    return d


@pytest.mark.parametrize("preset", list(o.PRESETS))
def test_oneshot_bit_exact_with_oracle(dfl, preset, pg11):
    """lib.rs:306-485 + tests/test.rs inputs; every stream equals the oracle's byte for byte."""
    opts = o.PRESETS[preset]()
    for name, data in _inputs(pg11).items():
        got = dfl.deflate_bytes_conf(data, _copts(dfl, opts))
        assert zlib.decompress(got, -15) == data, (name, preset)
        want = o.compress(data, opts, o.RAW)
        assert got == want, (name, preset, len(got), len(want))


@pytest.mark.parametrize("preset", ["default", "fast"])
def test_afl_inputs_zlib(dfl, preset):
    """tests/test.rs:138-161: 45 fuzzer-found inputs x {default, fast}, zlib container."""
    opts = o.PRESETS[preset]()
    for name in sorted(os.listdir(os.path.join(FIXTURES, "afl"))):
        data = fixture_bytes("afl/" + name)
        got = dfl.deflate_bytes_zlib_conf(data, _copts(dfl, opts))
        assert zlib.decompress(got) == data, name
        assert got == o.compress(data, opts, o.ZLIB), name


def test_pinned_sizes(dfl):
    """lib.rs:383-391 (5 bytes), tests/test.rs:58-64 (30 bytes), zlib.rs:69-86 (78 9C), empty input."""
    assert len(dfl.deflate_bytes(bytes([10, 10, 10, 10, 10, 55]))) == 5
    short = fixture_bytes("short.bin")
    z = dfl.deflate_bytes_zlib(short)
    assert len(z) == 30 and z[:2] == b"\x78\x9c" and zlib.decompress(z) == short
    assert dfl.deflate_bytes(b"") == b"\x03\x00"
    assert zlib.decompress(dfl.deflate_bytes_zlib(b"")) == b""


def test_issue_44_three_byte_values(dfl):
    """tests/test.rs:115-136: 25 MiB made of three distinct byte values."""
    data = zlib.decompress(fixture_bytes("issue_44.zlib"))
    for preset in ("default", "fast"):
        opts = o.PRESETS[preset]()
        got = dfl.deflate_bytes_zlib_conf(data, _copts(dfl, opts))
        assert zlib.decompress(got) == data
        assert got == o.compress(data, opts, o.ZLIB), preset


def test_entropy_stage_bit_exact_on_oracle_tokens(dfl, pg11):
    """huffman_lengths.rs:167-369 + encoder_state.rs:58-105 + bitstream.rs:76-106 in isolation: the
    block cutter / code builder / bit packer fed the oracle's tokens reproduces the oracle's bytes."""
    L = dfl._native.lib()
    for data in (pg11, bytes(70000), np.random.default_rng(3).integers(0, 256, 90000, dtype=np.uint8).tobytes()):
        for preset in ("default", "fast", "huffman_only"):
            opts = o.PRESETS[preset]()
            litlen, dist, _ = o.lz77_tokens(data, opts)
            toks = np.where(dist == 0, litlen, (litlen + 3) | (dist << 9)).astype(np.uint32)
            cap = L.dfl_bound(len(data), 0)
            out = ctypes.create_string_buffer(cap)
            n = ctypes.c_size_t()
            rc = L.dfl_encode_tokens(data, len(data), toks.ctypes.data_as(ctypes.c_void_p), len(toks), out, cap, ctypes.byref(n))
            assert rc == 0
            assert out.raw[: n.value] == o.compress(data, opts, o.RAW), preset


def test_lz77_stage_tokens_equal_oracle(dfl, pg11):
    """lz77.rs:305-547 + matching.rs:87-166: the token stream itself, not just its size."""
    L = dfl._native.lib()
    for preset in ("default", "fast", "high", "rle"):
        opts = o.PRESETS[preset]()
        litlen, dist, _ = o.lz77_tokens(pg11, opts)
        want = np.where(dist == 0, litlen, (litlen + 3) | (dist << 9)).astype(np.uint32)
        got = np.zeros(len(pg11) + 16, dtype=np.uint32)
        n = ctypes.c_size_t()
        co = dfl._native.dfl_options(opts.max_hash_checks, opts.lazy_if_less_than, opts.matching_type, 0)
        rc = L.dfl_lz77_tokens(pg11, len(pg11), ctypes.byref(co), got.ctypes.data_as(ctypes.c_void_p), len(got), ctypes.byref(n))
        assert rc == 0 and n.value == len(want)
        assert (got[: n.value] == want).all(), preset


def test_adler32_on_device(dfl, pg11):
    import torch
    L = dfl._native.lib()
    for data in (b"", b"a", pg11, bytes([255]) * 300000, np.random.default_rng(5).integers(0, 256, 1 << 20, dtype=np.uint8).tobytes()):
        t = torch.frombuffer(bytearray(data) if data else bytearray(1), dtype=torch.uint8).cuda()
        a = ctypes.c_uint32()
        assert L.dfl_adler32_device(ctypes.c_void_p(t.data_ptr()), len(data), ctypes.byref(a), None) == 0
        assert a.value == zlib.adler32(data)


def test_custom_options_bit_exact(dfl, pg11):
    data = pg11[:120000]
    for checks, lazy, mt in ((4, 8, 1), (16, 258, 1), (3, 40, 1), (64, 4, 1), (7, 0, 0), (300, 64, 1), (1, 3, 1)):
        opts = o.Options(checks, lazy, mt, 0)
        got = dfl.deflate_bytes_conf(data, _copts(dfl, opts))
        assert got == o.compress(data, opts, o.RAW), (checks, lazy, mt)


def test_device_api_and_overflow(dfl, pg11):
    import torch
    src = torch.frombuffer(bytearray(pg11), dtype=torch.uint8).cuda()
    out, n = dfl.compress_device(src, dfl.Compression.Default, dfl.ZLIB)
    got = bytes(out[:n].cpu().numpy())
    assert got == o.compress(pg11, o.opts_default(), o.ZLIB)
    small = torch.empty(1024, dtype=torch.uint8, device="cuda")
    with pytest.raises(dfl.DeflateB200Error) as ei:
        dfl.compress_device(src, dfl.Compression.Default, dfl.RAW, out=small)
    assert ei.value.status == -5   # DFL_E_OVERFLOW, nothing written past the buffer


def test_large_synthetic_roundtrip_and_ratio(dfl):
    """BASELINE config 2 shape at 64 MiB: round trip, and the size equals the oracle's on a prefix."""
    import datagen
    import torch
    data = datagen.silesia_mix(64 << 20)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    out, n = dfl.compress_device(src, dfl.Compression.Default, dfl.RAW)
    got = bytes(out[:n].cpu().numpy())
    assert zlib.decompress(got, -15) == data
    prefix = data[: 4 << 20]
    assert dfl.deflate_bytes(prefix) == o.compress(prefix, o.opts_default(), o.RAW)
    out, n = dfl.compress_device(src, dfl.Compression.Fast, dfl.ZLIB)
    assert zlib.decompress(bytes(out[:n].cpu().numpy())) == data


# ---------------------------------------------------------------- writers (writer.rs:502-660)
def test_writer_chunked_equals_oneshot(dfl, pg11):
    """lib.rs:408-433: streaming output == one-shot output for any chunking."""
    want = dfl.deflate_bytes_zlib(pg11)
    assert want == o.compress(pg11, o.opts_default(), o.ZLIB)
    for chunk in (50, 400, 32768, 65794, 50000):
        sink = bytearray()
        enc = dfl.write.ZlibEncoder(sink, dfl.Compression.Default)
        for i in range(0, len(pg11), chunk):
            enc.write_all(pg11[i:i + chunk])
        enc.finish()
        assert bytes(sink) == want, chunk


def test_writer_reset_is_deterministic(dfl, pg11):
    """writer.rs:537-568"""
    for cls in (dfl.write.DeflateEncoder, dfl.write.ZlibEncoder):
        enc = cls(bytearray(), dfl.CompressionOptions.default())
        enc.write_all(pg11)
        res1 = enc.reset(bytearray())
        enc.write_all(pg11)
        res2 = enc.finish()
        assert bytes(res1) == bytes(res2) and len(res1) > 0


def test_writer_sync_flush(dfl, pg11):
    """writer.rs:570-660, tests/test.rs:113-136"""
    sink = bytearray()
    enc = dfl.write.DeflateEncoder(sink, dfl.CompressionOptions.default())
    split = len(pg11) // 2
    enc.write_all(pg11[:split])
    enc.flush()
    enc.flush()
    assert bytes(sink[-4:]) == b"\x00\x00\xff\xff"
    enc.write_all(pg11[split:split + 2])
    enc.flush()
    enc.write_all(pg11[split + 2:])
    enc.finish()
    assert zlib.decompress(bytes(sink), -15) == pg11
    # the flushed stream matches the oracle's writer byte for byte as well
    s = o.Stream(o.opts_default(), o.RAW)
    s.write(pg11[:split]); s.flush(); s.flush(); s.write(pg11[split:split + 2]); s.flush(); s.write(pg11[split + 2:])
    assert bytes(sink) == s.finish()
    sink = bytearray()
    enc = dfl.write.DeflateEncoder(sink, dfl.CompressionOptions.default())
    enc.flush(); enc.write_all(bytes([1, 2])); enc.flush(); enc.write_all(bytes([3])); enc.flush()
    enc.finish()
    assert zlib.decompress(bytes(sink), -15) == bytes([1, 2, 3])


def test_writer_small_sink_and_checksum(dfl, pg11):
    """tests/test.rs:163-200 (a sink that takes <= 2 bytes per call) and writer.rs:248 checksum()."""
    class SmallWriter:
        def __init__(self):
            self.data = bytearray()

        def write(self, b):
            k = min(2, len(b))
            self.data += b[:k]
            return k

    w = SmallWriter()
    enc = dfl.write.ZlibEncoder(w, dfl.Compression.Fast)
    enc.write_all(pg11[:5000])
    enc.flush()
    assert enc.checksum() == zlib.adler32(pg11[:5000])
    enc.write_all(pg11[5000:9000])
    assert enc.checksum() == zlib.adler32(pg11[:9000])
    enc.finish()
    assert zlib.decompress(bytes(w.data)) == pg11[:9000]
