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


def test_lazy_below_three_bit_exact(dfl, pg11):
    """MatchingType::Lazy with lazy_if_less_than in {0, 1, 2} (0 is documented, compression_options.rs:93): the
    reference's length-2 results from spurious chain entries (matching.rs:161-165, chained_hash_table.rs:34-51)
    and its per-call re-derivation of ignore_next (lz77.rs:331) become visible in the output.  These option sets
    run the reference's loop itself on the device (k_lz77_seq), one-shot in every container and through a writer
    that is finished without intermediate flushes."""
    rng = np.random.default_rng(1)
    inputs = {"pg11": pg11, "issue_18": fixture_bytes("issue_18_201911.bin"),
              "random4": rng.integers(0, 4, 200000, dtype=np.uint8).tobytes(),
              "random16": rng.integers(0, 16, 200000, dtype=np.uint8).tobytes(), "zeros": bytes(100000), "empty": b"",
              "pg70000": pg11[:70000]}
    for checks, lazy, mt in ((128, 0, 1), (128, 1, 1), (128, 2, 1), (1, 0, 1), (4, 2, 1), (1768, 2, 1)):
        opts = o.Options(checks, lazy, mt, 0)
        for name, data in inputs.items():
            want = o.compress(data, opts, o.RAW)
            assert dfl.deflate_bytes_conf(data, _copts(dfl, opts)) == want, (checks, lazy, name)
    opts = o.Options(128, 0, 1, 0)
    assert dfl.deflate_bytes_zlib_conf(pg11, _copts(dfl, opts)) == o.compress(pg11, opts, o.ZLIB)
    enc = dfl.write.DeflateEncoder(bytearray(), _copts(dfl, opts))
    for i in range(0, len(pg11), 50000):
        enc.write_all(pg11[i:i + 50000])
    assert bytes(enc.finish()) == o.compress(pg11, opts, o.RAW)


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
    # the size reported with DFL_E_OVERFLOW is a capacity that suffices: retrying with exactly it succeeds
    L = dfl._native.lib()
    opts = dfl.CompressionOptions.default()._c()
    need = ctypes.c_size_t()
    rc = L.dfl_compress_device(ctypes.c_void_p(src.data_ptr()), src.numel(), ctypes.byref(opts), dfl.RAW, None, 0,
                               ctypes.c_void_p(small.data_ptr()), small.numel(), ctypes.byref(need), None)
    assert rc == -5 and need.value > small.numel()
    exact = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    rc = L.dfl_compress_device(ctypes.c_void_p(src.data_ptr()), src.numel(), ctypes.byref(opts), dfl.RAW, None, 0,
                               ctypes.c_void_p(exact.data_ptr()), exact.numel(), ctypes.byref(need), None)
    assert rc == 0 and bytes(exact[:need.value].cpu().numpy()) == o.compress(pg11, o.opts_default(), o.RAW)


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


# ---------------------------------------------------------------- gzip (lib.rs:241-286, writer.rs:331-467)
def test_crc32_on_device(dfl, pg11):
    import torch
    L = dfl._native.lib()
    rng = np.random.default_rng(7)
    cases = [b"", b"a", pg11, bytes([255]) * 300000, rng.integers(0, 256, (1 << 20) + 12345, dtype=np.uint8).tobytes(),
             rng.integers(0, 256, 65536, dtype=np.uint8).tobytes(), rng.integers(0, 256, 65537, dtype=np.uint8).tobytes()]
    for data in cases:
        t = torch.frombuffer(bytearray(data) if data else bytearray(1), dtype=torch.uint8).cuda()
        c = ctypes.c_uint32()
        assert L.dfl_crc32_device(ctypes.c_void_p(t.data_ptr()), len(data), ctypes.byref(c), None) == 0
        assert c.value == zlib.crc32(data), len(data)
        assert c.value == o.lib().dfo_crc32(0, data, len(data))
    # unaligned device pointer
    data = cases[4]
    t = torch.frombuffer(bytearray(b"xyz" + data), dtype=torch.uint8).cuda()
    c = ctypes.c_uint32()
    assert L.dfl_crc32_device(ctypes.c_void_p(t.data_ptr() + 3), len(data), ctypes.byref(c), None) == 0
    assert c.value == zlib.crc32(data)


def test_gzip_oneshot_equals_oracle(dfl, pg11):
    """lib.rs:393-406: gzip one-shot; default GzBuilder header, CRC-32 and ISIZE trailer."""
    for name, data in _inputs(pg11).items():
        for preset in ("default", "fast"):
            opts = o.PRESETS[preset]()
            got = dfl.deflate_bytes_gzip_conf(data, _copts(dfl, opts))
            assert zlib.decompress(got, 31) == data, (name, preset)
            assert got == o.compress(data, opts, o.GZIP), (name, preset)
    assert dfl.deflate_bytes_gzip(pg11) == o.compress(pg11, o.opts_default(), o.GZIP)


def test_gzip_builder_header_fields(dfl, pg11):
    """writer.rs:473-491 / lib.rs:393-406 assert the comment field survives; here the whole header does."""
    b = dfl.GzBuilder().comment(b"Test").filename(b"pg11.txt").extra(b"ab").mtime(1234567)
    got = dfl.deflate_bytes_gzip_conf(pg11, dfl.Compression.Default, b)
    hdr = b.into_header()
    assert got[:len(hdr)] == hdr and hdr[3] == 4 | 8 | 16 and b"Test\0" in hdr
    d = zlib.decompressobj(31)
    assert d.decompress(got) == pg11 and d.eof
    # the deflate stream and the trailer do not depend on the header
    ref = o.compress(pg11, o.opts_default(), o.GZIP)
    assert got[len(hdr):] == ref[10:]


def test_gz_encoder_writer(dfl, pg11):
    """writer.rs:331-467: chunked writes == one-shot; checksum(); reset_with_builder; sync flush."""
    want = dfl.deflate_bytes_gzip(pg11)
    for chunk in (50, 32768, 70001):
        sink = bytearray()
        enc = dfl.write.GzEncoder(sink, dfl.Compression.Default)
        for i in range(0, len(pg11), chunk):
            enc.write_all(pg11[i:i + chunk])
        assert enc.checksum() == zlib.crc32(pg11)
        enc.finish()
        assert bytes(sink) == want, chunk
    s = o.Stream(o.opts_default(), o.GZIP)
    s.write(pg11[:1000]); s.flush(); s.write(pg11[1000:])
    sink = bytearray()
    enc = dfl.write.GzEncoder(sink, dfl.CompressionOptions.default())
    enc.write_all(pg11[:1000]); enc.flush(); enc.write_all(pg11[1000:])
    enc.finish()
    assert bytes(sink) == s.finish()
    enc = dfl.write.GzEncoder.from_builder(dfl.GzBuilder().comment(b"one"), bytearray(), dfl.Compression.Fast)
    enc.write_all(pg11[:5000])
    first = enc.reset_with_builder(bytearray(), dfl.GzBuilder().comment(b"two"))
    enc.write_all(pg11[5000:12000])
    second = enc.finish()
    assert zlib.decompress(bytes(first), 31) == pg11[:5000] and b"one\0" in bytes(first[:20])
    assert zlib.decompress(bytes(second), 31) == pg11[5000:12000] and b"two\0" in bytes(second[:20])


# ---------------------------------------------------------------- match stage on every kind of input
def test_match_stage_bit_exact_on_all_kinds(dfl, pg11):
    """k_match (entry walk) + the parse stage's warp-cooperative resolution of long records implement
    matching.rs:87-166 over the sorted windows; the result must equal the oracle on every kind of input."""
    import datagen
    inputs = _inputs(pg11)
    inputs["mix2m"] = datagen.silesia_mix(2 << 20)
    inputs["enwik1m"] = datagen.enwik_like(1 << 20)
    inputs["png1m"] = datagen.png_idat_like(1 << 20)
    inputs["bin1m"] = datagen.binary_like(1 << 20)
    for name in sorted(os.listdir(os.path.join(FIXTURES, "afl")))[:12]:
        inputs["afl/" + name] = fixture_bytes("afl/" + name)
    for preset in ("default", "fast"):
        opts = o.PRESETS[preset]()
        for name, data in inputs.items():
            got = dfl.deflate_bytes_conf(data, _copts(dfl, opts))
            assert got == o.compress(data, opts, o.RAW), (name, preset)
    for checks, lazy, mt in ((4, 8, 1), (16, 258, 1), (64, 4, 1), (7, 0, 0), (128, 33, 1), (129, 32, 1), (100, 20, 0)):
        opts = o.Options(checks, lazy, mt, 0)
        data = inputs["mix2m"][:600000]
        assert dfl.deflate_bytes_conf(data, _copts(dfl, opts)) == o.compress(data, opts, o.RAW), (checks, lazy, mt)


def test_streaming_dictionary(dfl, pg11):
    """Pieces encoded with the previous 32 KiB as dictionary (begin > 0)."""
    s = o.Stream(o.opts_default(), o.ZLIB)
    sink = bytearray()
    enc = dfl.write.ZlibEncoder(sink, dfl.Compression.Default)
    for lo, hi in ((0, 40000), (40000, 40010), (40010, 120000), (120000, len(pg11))):
        s.write(pg11[lo:hi]); s.flush()
        enc.write_all(pg11[lo:hi]); enc.flush()
    enc.finish()
    assert bytes(sink) == s.finish()


# ---------------------------------------------------------------- pieces of one stream
def test_pieces_concatenate_to_the_flushed_reference_stream(dfl, pg11):
    """dfl_compress_device_piece: pieces encoded independently (as the ranks of a multi-GPU job do),
    each with the 32 KiB in front of it as dictionary, concatenate into the stream the reference's
    writer produces with flush() at the piece boundaries."""
    import datagen
    import torch
    from deflate_rs_b200 import sharding
    small = [(0, 36000)] + [(lo, min(len(pg11), lo + 8192)) for lo in range(36000, len(pg11), 8192)]
    cases = ((pg11, sharding.piece_bounds(len(pg11), 3, 4096)),
             (datagen.silesia_mix(3 << 20), sharding.piece_bounds(3 << 20, 4, 1 << 16)),
             (pg11, small))
    # Compression::Fast takes the one-candidate path (the sort kernel settles the matches; the candidate of the first
    # position of a bucket lies in the window in front, which for a piece may be dictionary only)
    for preset, oopts in ((dfl.Compression.Default, o.opts_default()), (dfl.Compression.Fast, o.opts_fast())):
        for data, bounds in cases:
            src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
            bounds = [b for b in bounds if b[1] > b[0]]
            s = o.Stream(oopts, o.RAW)
            got = b""
            for g, (lo, hi) in enumerate(bounds):
                last = g + 1 == len(bounds)
                out, n = sharding.encode_piece_device(src, lo, hi, preset, last)
                got += bytes(out[:n].cpu().numpy())
                s.write(data[lo:hi])
                if not last:
                    s.flush()
            assert got == s.finish(), (len(data), len(bounds), preset)
            assert zlib.decompress(got, -15) == data


def test_sync_flush_inside_first_window_is_valid_but_not_the_reference_quirk(dfl, pg11):
    """Known divergence (DESIGN.md "Parity statement"): when flush() is called at a stream offset <= 32768
    and the next bytes arrive in a separate write() call, the reference re-seeds its rolling hash with the
    first two bytes of the stream (lz77.rs:628-639 runs again because `add_initial` is a local of the
    earlier call, :606-615), so the two positions after the flush are hashed wrongly and lose their
    matches.  The oracle restates that; the GPU path hashes them correctly.  Both streams are valid."""
    sink = bytearray()
    enc = dfl.write.DeflateEncoder(sink, dfl.Compression.Default)
    s = o.Stream(o.opts_default(), o.RAW)
    for part in (pg11[:30000], pg11[30000:60000]):
        enc.write_all(part); enc.flush()
        s.write(part); s.flush()
    enc.finish()
    ref = s.finish()
    assert zlib.decompress(bytes(sink), -15) == pg11[:60000] and zlib.decompress(ref, -15) == pg11[:60000]
    assert len(sink) <= len(ref)            # the quirk costs the reference two literals here
    # outside the first window the two are identical again
    sink = bytearray()
    enc = dfl.write.DeflateEncoder(sink, dfl.Compression.Default)
    s = o.Stream(o.opts_default(), o.RAW)
    for part in (pg11[:32769], pg11[32769:60000]):
        enc.write_all(part); enc.flush()
        s.write(part); s.flush()
    enc.finish()
    assert bytes(sink) == s.finish()


def test_batch_of_independent_streams(dfl, pg11):
    """dfl_compress_device_batch (BASELINE config 4 shape): every stream equals the one-shot result."""
    import datagen
    import torch
    datas = [datagen.png_idat_like(200000 + 1000 * i, 0x1DA7 + i) for i in range(20)] + [b"", b"x", pg11]
    srcs = [torch.frombuffer(bytearray(d) if d else bytearray(1), dtype=torch.uint8).cuda()[:len(d)] for d in datas]
    for wrap, owrap, wbits in ((dfl.ZLIB, o.ZLIB, 15), (dfl.RAW, o.RAW, -15), (dfl.GZIP, o.GZIP, 31)):
        outs, sizes = dfl.compress_device_batch(srcs, dfl.Compression.Default, wrap)
        for d, out, s in zip(datas, outs, sizes):
            got = bytes(out[:s].cpu().numpy())
            assert zlib.decompress(got, wbits) == d
            assert got == o.compress(d, o.opts_default(), owrap), len(d)


def test_host_batch_of_independent_streams(dfl, pg11):
    """dfl_compress_batch: host buffers in and out, more members than pipelines in the pool (the lanes are reused),
    ragged sizes incl. empty; every member equals the reference algorithm's one-shot result."""
    import datagen
    datas = [datagen.png_idat_like(50000 + 7001 * i, 0xBA7C + i) for i in range(37)] + [b"", b"x", pg11, b""]
    for wrap, owrap, opts, oopts in ((dfl.ZLIB, o.ZLIB, dfl.Compression.Default, o.opts_default()),
                                     (dfl.RAW, o.RAW, dfl.Compression.Fast, o.opts_fast()),
                                     (dfl.GZIP, o.GZIP, dfl.Compression.Best, o.opts_high())):
        got = dfl.compress_batch(datas, opts, wrap)
        assert len(got) == len(datas)
        for d, g in zip(datas, got):
            assert g == o.compress(d, oopts, owrap), len(d)


def test_host_batch_reports_overflow_per_member(dfl):
    """A member whose output buffer is too small gets DFL_E_OVERFLOW and its needed size; the others are unaffected."""
    import ctypes
    import numpy as np
    L = dfl._native.lib()
    datas = [bytes(range(256)) * 40, os.urandom(5000), b"abc" * 1000]
    srcs = [np.frombuffer(d, dtype=np.uint8) for d in datas]
    caps = [L.dfl_bound(len(d), dfl.ZLIB) + 64 for d in datas]
    caps[1] = 100
    outs = [np.zeros(c, dtype=np.uint8) for c in caps]
    k = len(datas)
    opts = dfl.CompressionOptions.default()._c()
    out_len = (ctypes.c_size_t * k)()
    status = (ctypes.c_int * k)()
    rc = L.dfl_compress_batch(k, (ctypes.c_void_p * k)(*[s.ctypes.data for s in srcs]), (ctypes.c_size_t * k)(*map(len, datas)),
                              ctypes.byref(opts), dfl.ZLIB, (ctypes.c_void_p * k)(*[x.ctypes.data for x in outs]),
                              (ctypes.c_size_t * k)(*caps), out_len, status)
    assert rc == dfl._native.E_OVERFLOW and list(status) == [0, dfl._native.E_OVERFLOW, 0]
    want = [o.compress(d, o.opts_default(), o.ZLIB) for d in datas]
    assert out_len[1] == len(want[1])
    for i in (0, 2):
        assert bytes(outs[i][:out_len[i]]) == want[i]


def test_trim_releases_the_scratch_and_the_next_call_rebuilds_it(dfl, pg11):
    """dfl_trim (the reference drops its Vec when deflate_bytes returns, lib.rs:141-146; the library keeps its device
    scratch between calls because allocating it costs more than encoding with it, and gives it back on request)."""
    import datagen
    import torch
    data = datagen.silesia_mix(64 << 20)
    want = dfl.deflate_bytes(data)
    sink = bytearray()
    enc = dfl.write.ZlibEncoder(sink, dfl.Compression.Default)   # parks its resources in the handle pool when dropped
    enc.write_all(pg11); enc.finish(); del enc
    dfl.compress_batch([pg11, pg11[:50000]], dfl.Compression.Default, dfl.ZLIB)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    dfl.trim()
    free1, _ = torch.cuda.mem_get_info()
    assert free1 - free0 > 20 * len(data), (free0, free1)           # the 64 MiB context alone held about 29 bytes per input byte
    assert dfl.deflate_bytes(data) == want
    assert dfl.deflate_bytes(pg11) == o.compress(pg11, o.opts_default(), o.RAW)


# ---------------------------------------------------------------- BASELINE.json's full sizes
def _inflate_equals(comp: bytes, data: bytes, wbits: int) -> bool:
    d = zlib.decompressobj(wbits)
    pos, step = 0, 64 << 20
    view = memoryview(comp)
    off = 0
    while off < len(view):
        out = d.decompress(view[off:off + (8 << 20)])
        off += 8 << 20
        if out != data[pos:pos + len(out)]:
            return False
        pos += len(out)
    out = d.flush()
    return out == data[pos:pos + len(out)] and pos + len(out) == len(data) and d.eof


def test_full_size_default_raw_roundtrip(dfl):
    """Config 2 at its real size (1 GiB, Compression::Default, raw deflate): the stream inflates to the
    input and equals the oracle's, all 340 MB of it."""
    import datagen
    import torch
    data = datagen.silesia_mix(1 << 30)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    out, n = dfl.compress_device(src, dfl.Compression.Default, dfl.RAW)
    comp = bytes(out[:n].cpu().numpy())
    assert _inflate_equals(comp, data, -15)
    # the whole stream against the oracle (about a minute of oracle time on one host core)
    ref = o.compress(data, o.opts_default(), o.RAW)
    assert len(comp) == len(ref) and comp == ref


def test_full_size_fast_zlib_roundtrip(dfl):
    """Config 3 at its real size (1 GiB enwik-like, Compression::Fast, ZlibEncoder framing): zlib verifies the
    Adler-32 computed on the device; the device checksums equal CPython's on the whole GiB."""
    import datagen
    import torch
    data = datagen.enwik_like(1 << 30)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    out, n = dfl.compress_device(src, dfl.Compression.Fast, dfl.ZLIB)
    comp = bytes(out[:n].cpu().numpy())
    assert _inflate_equals(comp, data, 15)
    L = dfl._native.lib()
    a, c = ctypes.c_uint32(), ctypes.c_uint32()
    assert L.dfl_adler32_device(ctypes.c_void_p(src.data_ptr()), len(data), ctypes.byref(a), None) == 0
    assert L.dfl_crc32_device(ctypes.c_void_p(src.data_ptr()), len(data), ctypes.byref(c), None) == 0
    assert a.value == zlib.adler32(data) and c.value == zlib.crc32(data)
    assert int.from_bytes(comp[-4:], "big") == a.value
    assert comp == o.compress(data, o.opts_fast(), o.ZLIB)   # the whole stream against the oracle


def test_full_size_high_binary(dfl):
    """Config 5 at its real size (256 MiB synthetic binary, CompressionOptions::high(): 1768 checks, lazy < 128,
    quarter-budget searches): whole stream against the oracle."""
    import datagen
    import torch
    data = datagen.binary_like(256 << 20)
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    out, n = dfl.compress_device(src, dfl.CompressionOptions.high(), dfl.RAW)
    comp = bytes(out[:n].cpu().numpy())
    assert comp == o.compress(data, o.opts_high(), o.RAW)


def test_full_size_png_chunks_batch(dfl):
    """Config 4's unit of work at its real size: 4 MiB PNG-IDAT-like chunks, one zlib stream each, through the batch
    calls (device and host buffers); 64 of the 256 chunks are compared with the oracle byte for byte (each costs
    the oracle half a second), the batch of all 256 is checked for inflating to its inputs."""
    import datagen
    import torch
    chunk = 4 << 20
    datas = [datagen.png_idat_like(chunk, 0x1DA7 + i) for i in range(256)]
    srcs = [torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda() for d in datas]
    outs, sizes = dfl.compress_device_batch(srcs, dfl.Compression.Default, dfl.ZLIB)
    comps = [bytes(x[:s_].cpu().numpy()) for x, s_ in zip(outs, sizes)]
    for i in range(0, 256, 4):
        assert comps[i] == o.compress(datas[i], o.opts_default(), o.ZLIB), i
    for i in range(256):
        assert zlib.decompress(comps[i]) == datas[i], i
    host = dfl.compress_batch(datas[:32], dfl.Compression.Default, dfl.ZLIB)
    assert host == comps[:32]


def test_bounded_budget_is_within_three_percent_of_default(dfl):
    """The bounded matcher of bench.py --config c2b: the reference's own algorithm at a chain budget of 24.  Still
    byte-identical to the reference at these options, and its stream is within 3 % of Compression::Default's size
    on the headline input (north_star's size bound)."""
    import datagen
    data = datagen.silesia_mix(16 << 20)
    opts = o.Options(24, 32, 1, 0)
    got = dfl.deflate_bytes_conf(data, _copts(dfl, opts))
    assert got == o.compress(data, opts, o.RAW)
    base = len(o.compress(data, o.opts_default(), o.RAW))
    assert len(got) <= 1.03 * base, (len(got), base)


def test_single_pipeline_runs_of_4_gib_are_refused_not_truncated(dfl):
    """Positions are 32 bit inside one pipeline run.  dfl_compress / dfl_compress_device split longer inputs into
    pieces themselves; the entry points that are one run by definition must refuse instead of wrapping around."""
    import torch
    L = dfl._native.lib()
    src = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    out = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    opts = dfl.CompressionOptions.default()._c()
    sz = ctypes.c_size_t()
    rc = L.dfl_compress_device_piece(ctypes.c_void_p(src.data_ptr()), 5 << 30, 0, ctypes.byref(opts), 2,
                                     ctypes.c_void_p(out.data_ptr()), out.numel(), ctypes.byref(sz), None)
    assert rc == -7   # DFL_E_UNSUPPORTED


def test_writer_pieces_without_flush_equal_oneshot(dfl, pg11):
    """The streaming handle encodes buffered input by itself once enough has accumulated (open pieces: parser
    state, uncoded tokens and the incomplete output byte are carried over).  Without a flush the reference's
    stream has no seam (lib.rs:408-433), so whatever the piece size, the bytes must equal the one-shot result."""
    import datagen
    mix = datagen.silesia_mix(3 << 20)
    cases = [(pg11, dfl.Compression.Default, o.opts_default()), (pg11, dfl.Compression.Fast, o.opts_fast()),
             (mix, dfl.Compression.Default, o.opts_default()), (pg11, dfl.CompressionOptions.rle(), o.PRESETS["rle"]()),
             (pg11, dfl.CompressionOptions.huffman_only(), o.PRESETS["huffman_only"]()),
             (bytes(300000), dfl.Compression.Default, o.opts_default()),
             (pg11[:100000], dfl.CompressionOptions.high(), o.PRESETS["high"]())]
    for data, opts, oopts in cases:
        for cls, owrap in ((dfl.write.DeflateEncoder, o.RAW), (dfl.write.ZlibEncoder, o.ZLIB), (dfl.write.GzEncoder, o.GZIP)):
            want = o.compress(data, oopts, owrap)
            for piece, chunk in ((4096, 1000), (20000, 7777), (65536, 65536), (100000, 250000), (1 << 20, 300000)):
                sink = bytearray()
                enc = cls(sink, opts)
                enc.set_piece_bytes(piece)
                for i in range(0, len(data), chunk):
                    enc.write_all(data[i:i + chunk])
                enc.finish()
                assert bytes(sink) == want, (len(data), cls.__name__, piece, chunk, len(sink), len(want))


def test_pageable_and_pinned_host_memory_give_the_same_bytes(dfl):
    """Pageable buffers of 2 MiB and more travel through the library's pinned slot rings and copy threads, pinned
    ones are copied directly: same bytes either way, for sizes on and around the slot (1 MiB) and ring (32 slots)
    boundaries, one-shot and through the writer with large writes."""
    import datagen
    import torch
    L = dfl._native.lib()
    opts = dfl.CompressionOptions.fast()._c()
    whole = datagen.silesia_mix((35 << 20) + 77, 0x57A6E)
    for n in ((2 << 20) - 1, 2 << 20, (2 << 20) + 1, (3 << 20) + 12345, 32 << 20, (33 << 20) + 5, len(whole)):
        data = whole[:n]
        want = o.compress(data, o.opts_fast(), o.ZLIB)
        cap = L.dfl_bound(n, dfl.ZLIB) + 64
        got = []
        for pinned in (False, True):
            src = torch.frombuffer(bytearray(data), dtype=torch.uint8)
            dst = torch.empty(cap, dtype=torch.uint8)
            if pinned:
                src, dst = src.pin_memory(), dst.pin_memory()
            m = ctypes.c_size_t()
            rc = L.dfl_compress(ctypes.c_void_p(src.data_ptr()), n, ctypes.byref(opts), dfl.ZLIB, None, 0,
                                ctypes.c_void_p(dst.data_ptr()), cap, ctypes.byref(m))
            assert rc == 0
            got.append(bytes(dst[:m.value].numpy()))
        assert got[0] == want and got[1] == want, n
    # the writer: writes large enough for the staged path, pieces small enough for several of them, both memory kinds
    want = o.compress(whole, o.opts_fast(), o.ZLIB)
    for pinned in (False, True):
        src = torch.frombuffer(bytearray(whole), dtype=torch.uint8)
        if pinned:
            src = src.pin_memory()
        e = L.dfl_encoder_new(ctypes.byref(opts), dfl.ZLIB, None, 0)
        assert L.dfl_encoder_set_piece_bytes(e, 9 << 20) == 0
        out = bytearray()
        p, ln = ctypes.POINTER(ctypes.c_uint8)(), ctypes.c_size_t()
        step = (5 << 20) + 3
        for off in range(0, len(whole), step):
            assert L.dfl_encoder_write(e, ctypes.c_void_p(src.data_ptr() + off), min(step, len(whole) - off), None) == 0
            L.dfl_encoder_take_output(e, ctypes.byref(p), ctypes.byref(ln))
            out += ctypes.string_at(p, ln.value)
            L.dfl_encoder_advance_output(e, ln.value)
        assert L.dfl_encoder_flush(e, dfl._native.FLUSH_FINISH) == 0
        L.dfl_encoder_take_output(e, ctypes.byref(p), ctypes.byref(ln))
        out += ctypes.string_at(p, ln.value)
        L.dfl_encoder_free(e)
        assert bytes(out) == want, pinned


def test_writers_on_several_threads(dfl, pg11):
    """Handles are independent: four threads stream at the same time (they share the copy threads and the pool of
    parked handle resources), several handles each; every stream equals the one-shot result."""
    import threading
    import datagen
    datas = [datagen.silesia_mix((5 << 20) + 1000 * i, 0x7EAD + i) for i in range(4)]
    wants = [o.compress(d, o.opts_fast(), o.ZLIB) for d in datas]
    errors = []

    def work(i):
        try:
            for rep in range(3):
                sink = bytearray()
                enc = dfl.write.ZlibEncoder(sink, dfl.Compression.Fast)
                enc.set_piece_bytes(1 << 20)
                step = (2 << 20) + 17 if rep else 70000
                for off in range(0, len(datas[i]), step):
                    enc.write_all(datas[i][off:off + step])
                enc.finish()
                if bytes(sink) != wants[i]:
                    errors.append((i, rep, len(sink), len(wants[i])))
        except Exception as ex:   # noqa: BLE001
            errors.append((i, repr(ex)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_writer_pieces_with_flushes_equal_the_reference_writer(dfl, pg11):
    """Open pieces and explicit sync flushes mixed (all flushes beyond the first window, see the divergence test)."""
    s = o.Stream(o.opts_default(), o.ZLIB)
    sink = bytearray()
    enc = dfl.write.ZlibEncoder(sink, dfl.Compression.Default)
    enc.set_piece_bytes(8192)
    for lo, hi, fl in ((0, 50000, True), (50000, 50010, False), (50010, 120000, True), (120000, len(pg11), False)):
        s.write(pg11[lo:hi]); enc.write_all(pg11[lo:hi])
        if fl:
            s.flush(); enc.flush()
    enc.finish()
    assert bytes(sink) == s.finish()


def test_writer_stream_longer_than_4_gib(dfl, pg11):
    """A single stream of 4.5 GiB through write::ZlibEncoder: pieces are encoded as the input arrives (positions are
    32 bit per piece only), the host buffer stays bounded, zlib inflates the result to the input and accepts the
    Adler-32 folded over all pieces."""
    block = (pg11 * 7)[: 1 << 20]
    total_blocks = 4608                      # 4.5 GiB
    class Sink:
        def __init__(self):
            self.d = zlib.decompressobj(15)
            self.pos = 0
            self.ok = True
            self.n_in = 0
        def write(self, b):
            self.n_in += len(b)
            out = self.d.decompress(bytes(b))
            # the input is periodic with period len(block): compare against the right rotation
            off = self.pos % len(block)
            k = 0
            while k < len(out):
                take = min(len(out) - k, len(block) - off)
                if out[k:k + take] != block[off:off + take]:
                    self.ok = False
                k += take
                off = 0
            self.pos += len(out)
            return len(b)
    sink = Sink()
    enc = dfl.write.ZlibEncoder(sink, dfl.Compression.Fast)
    for _ in range(total_blocks):
        enc.write_all(block)
    enc.finish()
    assert sink.ok and sink.pos == total_blocks * len(block) and sink.d.eof
    assert sink.n_in < sink.pos // 2


def test_oneshot_beyond_the_position_limit_goes_through_pieces(pg11):
    """dfl_compress routes inputs too long for one pipeline run through the streaming handle's pieces; with the
    switch-over point lowered (DFL_ONESHOT_PIECE_LIMIT) the same code path runs on small inputs, in a fresh
    process, and must give the ordinary bytes."""
    import subprocess
    import sys
    code = (
        "import sys, zlib; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import deflate_rs_b200 as d, oracle_lib as o\n"
        "data = open(%r, 'rb').read()\n"
        "for f, w in ((d.deflate_bytes, o.RAW), (d.deflate_bytes_zlib, o.ZLIB), (d.deflate_bytes_gzip, o.GZIP)):\n"
        "    assert f(data) == o.compress(data, o.opts_default(), w)\n"
        "assert d.deflate_bytes(b'abc') == o.compress(b'abc', o.opts_default(), o.RAW)\n"
        "import torch, datagen\n"
        "for blob in (data, datagen.silesia_mix(3 << 20), bytes(100000)):\n"
        "    src = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()\n"
        "    for w, ow in ((d.RAW, o.RAW), (d.ZLIB, o.ZLIB), (d.GZIP, o.GZIP)):\n"
        "        for opts, oo in ((d.Compression.Default, o.opts_default()), (d.Compression.Fast, o.opts_fast())):\n"
        "            out, n = d.compress_device(src, opts, w)\n"
        "            assert bytes(out[:n].cpu().numpy()) == o.compress(blob, oo, ow), (len(blob), w)\n"
        "print('ok')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
         os.path.join(FIXTURES, "pg11.txt"))
    env = dict(os.environ, DFL_ONESHOT_PIECE_LIMIT="20000")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_device_call_longer_than_4_gib(dfl, pg11):
    """dfl_compress_device on a 4.5 GiB device buffer: encoded as 1 GiB open pieces straight from HBM, one gzip
    member out; zlib inflates it to the input and accepts CRC-32 and ISIZE (= length mod 2^32)."""
    import torch
    block = (pg11 * 7)[: 1 << 20]
    total_blocks = 4608
    src = torch.frombuffer(bytearray(block), dtype=torch.uint8).cuda().repeat(total_blocks)
    assert src.numel() == total_blocks << 20
    out, n = dfl.compress_device(src, dfl.Compression.Fast, dfl.GZIP)
    comp = out[:n].cpu().numpy().tobytes()
    del src, out
    d = zlib.decompressobj(31)
    pos = 0
    ok = True
    view = memoryview(comp)
    for off in range(0, len(view), 4 << 20):
        got = d.decompress(view[off:off + (4 << 20)])
        k = 0
        o0 = pos % len(block)
        while k < len(got):
            take = min(len(got) - k, len(block) - o0)
            ok = ok and got[k:k + take] == block[o0:o0 + take]
            k += take
            o0 = 0
        pos += len(got)
    assert ok and pos == total_blocks << 20 and d.eof
    assert int.from_bytes(comp[-4:], "little") == (total_blocks << 20) & 0xFFFFFFFF


@pytest.mark.parametrize("preset", list(o.PRESETS))
def test_every_preset_on_every_kind_of_synthetic_data(dfl, preset):
    """4 MiB of each generator of the BASELINE configs (tests/datagen.py) at every preset, against the oracle:
    covers the quarter-budget walk (high), RLE and Huffman-only on data with long matches, many repairs of the
    speculative parse (sparse, PNG-like) and stored / fixed block choices (random)."""
    import datagen
    opts = o.PRESETS[preset]()
    for gen in (datagen.silesia_mix, datagen.enwik_like, datagen.png_idat_like, datagen.binary_like):
        data = gen(4 << 20)
        got = dfl.deflate_bytes_conf(data, _copts(dfl, opts))
        assert got == o.compress(data, opts, o.RAW), (preset, gen.__name__, len(got))
