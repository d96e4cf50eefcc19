"""Pins oracle/deflate_oracle.c against the reference's own known-answer tests.

Every test names the reference test it restates (file:line under /root/reference/src or tests/).
Inflation is always done by an independent decoder (CPython's zlib / libz), as the reference does
with miniz_oxide (src/test_utils.rs:23-27,70-72).
"""
import ctypes
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import oracle_lib as o
from conftest import FIXTURES, fixture_bytes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
L = o.lib()
U8P = ctypes.POINTER(ctypes.c_uint8)


def _u8(buf):
    return (ctypes.c_uint8 * len(buf))(*buf)


def _u16(buf):
    return (ctypes.c_uint16 * len(buf))(*buf)


def huffman_lengths(freqs, max_len):
    lens = (ctypes.c_uint8 * len(freqs))()
    L.dfo_huffman_lengths(_u16(freqs), len(freqs), max_len, lens)
    return list(lens)


def encode_lengths(lens):
    sym = (ctypes.c_uint8 * (len(lens) + 8))()
    arg = (ctypes.c_uint8 * (len(lens) + 8))()
    fr = (ctypes.c_uint16 * 19)()
    n = L.dfo_encode_lengths(_u8(lens), len(lens), sym, arg, fr)
    return [(sym[i], arg[i]) for i in range(n)], list(fr)


def lit(v):
    return (v, 0)


def zero(r):  # length_encode.rs:426-432
    if r <= 1:
        return (0, 0)
    return (17, r) if r <= 10 else (18, r)


def copy(r):
    return (16, r)


# ---------------------------------------------------------------- bitstream.rs:131-178
def test_bitwriter_known_answer():
    inp = [(3, 3), (10, 8), (88, 7), (0, 2), (0, 5), (0, 0), (238, 8), (126, 8), (161, 8), (10, 8),
           (238, 8), (174, 8), (126, 8), (174, 8), (65, 8), (142, 8), (62, 8), (10, 8), (1, 8), (161, 8),
           (78, 8), (62, 8), (158, 8), (206, 8), (10, 8), (64, 7), (0, 0), (24, 5), (0, 0), (174, 8),
           (126, 8), (193, 8), (174, 8)]
    expected = [83, 192, 2, 220, 253, 66, 21, 220, 93, 253, 92, 131, 28, 125, 20, 2, 66, 157, 124, 60,
                157, 21, 128, 216, 213, 47, 216, 21]
    out = (ctypes.c_uint8 * 64)()
    n = L.dfo_bitwriter_kat(_u16([v for v, _ in inp]), _u8([b for _, b in inp]), len(inp), out, 64)
    assert list(out[:n]) == expected


# ---------------------------------------------------------------- bit_reverse.rs:16-21
def test_reverse_bits():
    assert L.dfo_reverse_bits(0b0111_0100_0000_0000, 16) == 0b0000_0000_0010_1110
    assert L.dfo_reverse_bits(0b1100_1100_1100_1100, 16) == 0b0011_0011_0011_0011
    assert L.dfo_reverse_bits(0b11, 2) == 0b11
    assert L.dfo_reverse_bits(0b100, 3) == 0b001


# ---------------------------------------------------------------- huffman_table.rs constant tables
def test_tables_match_reference_source():
    ref = json.load(open(os.path.join(GOLDEN, "ref_tables.json")))
    eb, ev = ctypes.c_uint(), ctypes.c_uint()
    for stored in range(256):  # LENGTH_CODE / BASE_LENGTH / LENGTH_EXTRA_BITS_LENGTH
        code = L.dfo_length_code(stored + 3, ctypes.byref(eb), ctypes.byref(ev))
        n = ref["LENGTH_CODE"][stored]
        assert code == 257 + n
        assert eb.value == ref["LENGTH_EXTRA_BITS_LENGTH"][n]
        assert ev.value == stored - ref["BASE_LENGTH"][n]
    for dist in range(1, 32769):  # DISTANCE_CODES / DISTANCE_BASE / DISTANCE_EXTRA_BITS
        code = L.dfo_distance_code(dist, ctypes.byref(eb), ctypes.byref(ev))
        want = ref["DISTANCE_CODES"][dist - 1] if dist <= 256 else ref["DISTANCE_CODES"][256 + ((dist - 1) >> 7)]
        assert code == want
        assert eb.value == ref["DISTANCE_EXTRA_BITS"][code]
        assert ev.value == dist - (ref["DISTANCE_BASE"][code] + 1)
    # fixed code lengths, huffman_table.rs:32-42
    codes = (ctypes.c_uint16 * 288)()
    L.dfo_create_codes(_u8(ref["FIXED_CODE_LENGTHS"]), 288, codes)
    # huffman_table.rs:506-527 make_table_fixed
    assert codes[0] == 0b00001100 and codes[143] == 0b11111101 and codes[144] == 0b000010011
    assert codes[255] == 0b111111111 and codes[256] == 0 and codes[279] == 0b1110100
    assert codes[280] == 0b00000011 and codes[287] == 0b11100011
    dcodes = (ctypes.c_uint16 * 32)()
    L.dfo_create_codes(_u8([5] * 32), 32, dcodes)
    assert dcodes[0] == 0 and dcodes[5] == 20
    # (len 4, dist 5): length code 258 -> 0b00100000, distance code 4 -> 0b00100, 1 extra bit = 0
    assert L.dfo_length_code(4, None, None) == 258 and codes[258] == 0b00100000
    assert L.dfo_distance_code(5, ctypes.byref(eb), ctypes.byref(ev)) == 4
    assert dcodes[4] == 0b00100 and eb.value == 1 and ev.value == 0


def test_length_and_distance_code_kats():
    eb, ev = ctypes.c_uint(), ctypes.c_uint()
    # huffman_table.rs:439-459
    assert (L.dfo_length_code(4, ctypes.byref(eb), ctypes.byref(ev)), eb.value, ev.value) == (258, 0, 0)
    assert (L.dfo_length_code(165, ctypes.byref(eb), ctypes.byref(ev)), eb.value, ev.value) == (282, 5, 2)
    assert (L.dfo_length_code(257, ctypes.byref(eb), ctypes.byref(ev)), eb.value, ev.value) == (284, 5, 30)
    assert (L.dfo_length_code(258, ctypes.byref(eb), ctypes.byref(ev)), eb.value) == (285, 0)
    # huffman_table.rs:461-471
    for d, c in [(1, 0), (0, 0), (50000, 0), (6146, 25), (256, 15), (4733, 24), (257, 16)]:
        assert L.dfo_distance_code(d, None, None) == c
    # huffman_table.rs:473-485
    assert (L.dfo_distance_code(527, ctypes.byref(eb), ctypes.byref(ev)), eb.value, ev.value) == (18, 8, 0b1110)
    assert (L.dfo_distance_code(256, ctypes.byref(eb), None), eb.value) == (15, 6)
    assert (L.dfo_distance_code(4733, ctypes.byref(eb), None), eb.value) == (24, 11)


# ---------------------------------------------------------------- huffman_lengths.rs:374-384
def test_stored_padding():
    assert [L.dfo_stored_padding(i) for i in range(8)] == [5, 4, 3, 2, 1, 0, 7, 6]


# ---------------------------------------------------------------- length_encode.rs:569-660
def test_lengths_from_frequencies():
    assert huffman_lengths([1, 1, 5, 7, 10, 14], 4) == [4, 4, 3, 2, 2, 2]
    assert huffman_lengths([1, 5, 1, 7, 10, 14], 4) == [4, 3, 4, 2, 2, 2]
    res = huffman_lengths([0, 25, 0, 10, 2, 4], 4)
    assert res[0] == 0 and res[2] == 0 and res[1] < 4
    assert huffman_lengths([0, 0, 0, 0, 0, 0, 0, 0, 55, 0, 0, 0], 5) == [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0]
    assert huffman_lengths([0] * 30, 5) == [0] * 30
    freqs = [3] * 286
    freqs[55] = freqs[125] = 65535 // 3
    res = huffman_lengths(freqs, 15)
    assert res[55] < 3 and res[125] < 3
    # Kraft equality holds after the limiter
    assert sum(2 ** (15 - l) for l in res if l) == 2 ** 15


OPTIMAL_FREQS = (
    [0] * 10 + [44] + [0] * 21 + [68, 0, 14, 0, 0, 0, 0, 3, 7, 6, 1, 0, 12, 14, 9, 2, 6, 9, 4, 1, 1, 4, 1, 1, 0,
                                  0, 1, 3, 0, 6, 0, 0, 0, 4, 4, 1, 2, 5, 3, 2, 2, 9, 0, 0, 3, 1, 5, 5, 8, 0, 6, 10, 5, 2,
                                  0, 0, 1, 2, 0, 8, 11, 4, 0, 1, 3, 31, 13, 23, 22, 56, 22, 8, 11, 43, 0, 7, 33, 15, 45,
                                  40, 16, 1, 28, 37, 35, 26, 3, 7, 11, 9, 1, 1, 0, 1] + [0] * 130 +
    [1, 126, 114, 66, 31, 41, 25, 15, 21, 20, 16, 15, 10, 7, 5, 1, 1])


def test_optimal_lengths_7701_bits():
    """length_encode.rs:606-660: same total as miniz for this table."""
    assert len(OPTIMAL_FREQS) == 273
    lens = huffman_lengths(OPTIMAL_FREQS, 15)
    assert sum(f * l for f, l in zip(OPTIMAL_FREQS, lens)) == 7701


# ---------------------------------------------------------------- length_encode.rs:440-567
def test_encode_lengths_vectors():
    ref = json.load(open(os.path.join(GOLDEN, "ref_tables.json")))
    _, fr = encode_lengths(ref["FIXED_CODE_LENGTHS"])
    assert fr[0:7] == [0] * 7 and fr[10:16] == [0] * 6 and fr[17:19] == [0, 0]

    enc, _ = encode_lengths([0, 0, 5, 0, 15, 1, 0, 0, 0, 2, 4, 4, 4, 4, 3, 5, 5, 5, 5])
    assert enc == [lit(0), lit(0), lit(5), lit(0), lit(15), lit(1), zero(3), lit(2), lit(4), copy(3), lit(3),
                   lit(5), copy(3)]
    enc, _ = encode_lengths([0, 0, 0, 5, 2, 3, 0, 0, 0])
    assert enc == [zero(3), lit(5), lit(2), lit(3), zero(3)]
    enc, _ = encode_lengths([0, 0, 0, 3, 3, 3, 5, 4, 4, 4, 4, 0, 0])
    assert enc == [zero(3), lit(3), lit(3), lit(3), lit(5), lit(4), copy(3), lit(0), lit(0)]

    lens = ([0] * 10 + [9, 0, 0, 9] + [0] * 18 + [6, 0, 0, 0, 8, 0, 0, 0, 0, 8, 0, 0, 7, 8, 7, 8, 6, 6, 8, 0, 7, 6,
                                                  7, 8, 7, 7, 8, 0, 0, 0, 0, 0, 8, 8, 0, 8, 7, 0, 10, 8, 0, 8, 0, 10,
                                                  10, 8, 8, 10, 8, 0, 8, 7, 0, 10, 0, 7] + [0] * 9 +
            [6, 7, 7, 7, 6, 7, 8, 8, 6, 0, 0, 8, 8, 7, 8, 8, 0, 7, 6, 6, 8, 8, 8, 10, 10] + [0] * 133 +
            [10, 4, 3, 3, 4, 4, 5, 5, 5, 5, 5, 8, 8, 6, 7, 8, 10, 10, 0, 9] +
            [0, 0, 0, 0, 0, 0, 0, 8, 8, 8, 8, 6, 6, 5, 5, 5, 5, 6, 5, 5, 4, 4, 4, 4, 4, 4, 3, 4, 3, 4])
    enc, _ = encode_lengths(lens)
    assert enc[:10] == [zero(10), lit(9), lit(0), lit(0), lit(9), zero(18), lit(6), zero(3), lit(8), zero(4)]
    assert enc[10:20] == [lit(8), lit(0), lit(0), lit(7), lit(8), lit(7), lit(8), lit(6), lit(6), lit(8)]

    assert encode_lengths([1, 1, 1, 2])[0] == [lit(1), lit(1), lit(1), lit(2)]
    assert encode_lengths([0, 0, 3])[0] == [lit(0), lit(0), lit(3)]
    assert encode_lengths([0, 0, 0, 5, 2])[0] == [zero(3), lit(5), lit(2)]
    assert encode_lengths([0, 0, 0, 5, 0])[0][-1] != lit(5)
    assert encode_lengths([0, 4, 4, 4, 4, 0])[0][-1] == zero(0)


def _expand_rle(enc):
    out = []
    for s, a in enc:
        if s <= 15:
            out.append(s)
        elif s == 16:
            out.extend([out[-1]] * a)
        else:
            out.extend([0] * a)
    return out


def test_encode_lengths_roundtrip_random():
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(1, 320))
        lens = [int(x) for x in rng.choice([0, 0, 0, 3, 4, 4, 5, 8, 15], size=n)]
        enc, fr = encode_lengths(lens)
        assert _expand_rle(enc) == lens
        assert sum(fr) == len(enc)
        for s, a in enc:
            assert (s <= 15) or (s == 16 and 3 <= a <= 6) or (s == 17 and 3 <= a <= 10) or (s == 18 and 11 <= a <= 138)


# ---------------------------------------------------------------- matching.rs:297-343
def test_match_length():
    arr = bytes([5, 5, 5, 5, 5, 9, 9, 2, 3, 5, 5, 5, 5, 5])
    assert L.dfo_get_match_length(arr, len(arr), 9, 0) == 5
    assert L.dfo_get_match_length(arr, len(arr), 9, 7) == 0
    assert L.dfo_get_match_length(arr, len(arr), 10, 0) == 4


def _longest(data, fill_len, pos, prev_len, checks):
    ln, ds = ctypes.c_size_t(), ctypes.c_size_t()
    L.dfo_longest_match_filled(data, len(data), fill_len, pos, prev_len, checks, ctypes.byref(ln), ctypes.byref(ds))
    return ln.value, ds.value


def test_longest_match_kats():
    # get_longest_match (matching.rs:310-327): longest_match_current = position data.len()-? with the
    # table filled up to and including the searched position (hash needs 2 bytes look-ahead).
    data = b"xTest data, Test_data,zTest data"
    assert _longest(data, 23 + 1 + 3 - 1, 23, 0, 4096) == (9, 22)
    arr2 = bytes([10, 10, 10, 10, 10, 10, 10, 10, 2, 3, 5, 10, 10, 10, 10, 10])
    # the table holds positions 0..4; the head is position 4 whose nearest candidate is 3
    assert _longest(arr2, 3 + 1 + 1 + 2, 4, 0, 4096) == (4, 1)
    # match_index_zero (matching.rs:331-343)
    assert _longest(b"AAAAAAA", 5, 1, 0, 4096) == (6, 1)


# ---------------------------------------------------------------- chained_hash_table.rs:274-350
def test_hash_table_invariants():
    head = (ctypes.c_uint16 * 32768)()
    prev = (ctypes.c_uint16 * 32768)()
    L.dfo_hash_table_filled(b"", 0, head, prev)  # initial_chains
    assert list(head) == list(range(32768)) and list(prev) == list(range(32768))
    data = bytes(range(0, 255))  # table_unique: (255u8..0) is an empty range in Rust
    L.dfo_hash_table_filled(data, len(data), head, prev)
    h = 0
    for b in data:
        h = ((h << 5) ^ b) & 0x7FFF
    current_head = head[h]
    assert prev[current_head & 0x7FFF] == h


# ---------------------------------------------------------------- lz77.rs:938-1033
def _lz_decode(litlen, dist):
    out = bytearray()
    for ll, d in zip(litlen.tolist(), dist.tolist()):
        if d == 0:
            out.append(ll)
        else:
            for _ in range(ll + 3):
                out.append(out[-d])
    return bytes(out)


def test_lz77_token_kats(pg11):
    high = o.opts_high()
    ll, d, _ = o.lz77_tokens(b"Deflate late", high)  # compress_short
    assert _lz_decode(ll, d) == b"Deflate late" and (ll[-1] + 3, d[-1]) == (4, 5)
    ll, d, _ = o.lz77_tokens(b"nba badger nbadger", high)  # lazy
    assert d[-1] != 0 and ll[-1] + 3 == 6
    ll, d, _ = o.lz77_tokens(pg11, high)  # compress_long
    assert len(ll) < len(pg11) and _lz_decode(ll, d) == pg11
    for data in (bytes(32768), bytes(32768) + bytes([22]) * 32768,
                 bytes(32768) + bytes([22]) * 32768 + bytes([55]) * 32768,  # exact_window_size
                 bytes([35]) * 32768 + b"Test",  # border
                 bytes(2 * 32768 + 50) + b"\x01"):  # border_multiple_blocks
        ll, d, _ = o.lz77_tokens(data, high)
        assert _lz_decode(ll, d) == data and len(ll) < len(data)


def test_lz77_buffer_full_blocks(pg11):
    """lz77.rs:1126-1163 buffer_test_literals: a literal-only block is exactly 31744 tokens/bytes."""
    ll, d, ends = o.lz77_tokens(pg11, o.opts_huffman_only())
    assert (d == 0).all() and len(ll) == len(pg11)
    assert ends[0] == 31744 and ends[1] == 2 * 31744
    assert bytes(ll.astype("uint8")) == pg11
    for preset in ("default", "fast", "high", "rle"):
        ll, d, ends = o.lz77_tokens(pg11, o.PRESETS[preset]())
        assert _lz_decode(ll, d) == pg11
        assert all(e - s == 31744 for s, e in zip([0] + ends[:-2], ends[:-1]))
        assert len(ends) == len(ll) // 31744 + 1


# ---------------------------------------------------------------- compress.rs:333-345
def test_fixed_string_known_answer():
    want = bytes([0x73, 0x49, 0x4d, 0xcb, 0x49, 0x2c, 0x49, 0x55, 0x00, 0x11, 0x00])
    got = o.compress_fixed(b"Deflate late")
    assert got == want and zlib.decompress(got, -15) == b"Deflate late"


# ---------------------------------------------------------------- zlib.rs:69-86, lib.rs:383-391, tests/test.rs:58-64
def test_pinned_sizes_and_headers():
    z = o.compress(b"abc", o.opts_default(), o.ZLIB)
    assert z[:2] == b"\x78\x9c" and (z[0] * 256 + z[1]) % 31 == 0
    assert len(o.compress(bytes([10, 10, 10, 10, 10, 55]))) == 5
    short = fixture_bytes("short.bin")
    z = o.compress(short, o.opts_default(), o.ZLIB)
    assert len(z) == 30 and zlib.decompress(z) == short
    assert o.compress(b"") == b"\x03\x00"  # empty raw stream


# ---------------------------------------------------------------- lib.rs:306-485, tests/test.rs
@pytest.mark.parametrize("preset", list(o.PRESETS))
def test_roundtrip_fixtures(preset, pg11):
    opts = o.PRESETS[preset]()
    datas = [pg11, fixture_bytes("short.bin"), fixture_bytes("issue_18_201911.bin"), fixture_bytes("dump.bin"),
             b"", b"\x01", bytes([5, 6, 7, 8]), bytes(65537), bytes(61000), bytes([22]) * 32768 + bytes([5, 2, 55, 11, 12]),
             bytes([5]) * 100000, b"                    GNU GENERAL PUBLIC LICENSE"]
    for data in datas:
        assert zlib.decompress(o.compress(data, opts, o.RAW), -15) == data
        assert zlib.decompress(o.compress(data, opts, o.ZLIB)) == data
        assert zlib.decompress(o.compress(data, opts, o.GZIP), 31) == data


@pytest.mark.parametrize("preset", ["default", "fast"])
def test_roundtrip_afl_and_issue44(preset):
    """tests/test.rs:138-161 (afl crash inputs) and :115-136 (issue 44, 25 MiB of 3 byte values)."""
    opts = o.PRESETS[preset]()
    for name in sorted(os.listdir(os.path.join(FIXTURES, "afl"))):
        data = fixture_bytes("afl/" + name)
        assert zlib.decompress(o.compress(data, opts, o.ZLIB)) == data
    data = zlib.decompress(fixture_bytes("issue_44.zlib"))
    assert len(data) == 26214400
    assert zlib.decompress(o.compress(data, opts, o.ZLIB)) == data


def test_oracle_pins_unchanged():
    """The oracle's output for the reference fixtures is frozen in tests/golden/oracle_pins.json."""
    pins = json.load(open(os.path.join(GOLDEN, "oracle_pins.json")))
    for key in ["pg11.txt:default", "pg11.txt:fast", "pg11.txt:high", "pg11.txt:rle", "short.bin:default",
                "issue_18_201911.bin:default", "dump.bin:fast", "afl/" + sorted(os.listdir(os.path.join(FIXTURES, "afl")))[0] + ":default"]:
        f, preset = key.rsplit(":", 1)
        c = o.compress(fixture_bytes(f), o.PRESETS[preset](), o.RAW)
        assert [len(c), hashlib.sha256(c).hexdigest()] == pins[key], key


# ---------------------------------------------------------------- writer.rs:502-660, lib.rs:408-433
def test_writer_equals_oneshot_for_any_chunking(pg11):
    for preset in ("default", "fast", "rle"):
        opts = o.PRESETS[preset]()
        want = o.compress(pg11, opts, o.ZLIB)
        for chunk in (1 if preset == "fast" else 7, 50, 400, 32768, 65794, 50000):
            s = o.Stream(opts, o.ZLIB)
            for i in range(0, len(pg11), chunk):
                s.write(pg11[i:i + chunk])
            assert s.finish() == want, (preset, chunk)


def test_writer_reset_is_deterministic(pg11):
    for wrap in (o.RAW, o.ZLIB):
        s = o.Stream(o.opts_default(), wrap)
        s.write(pg11)
        res1 = s.reset()
        s.write(pg11)
        assert s.finish() == res1


def test_writer_sync_flush(pg11):
    s = o.Stream(o.opts_default(), o.RAW)
    split = len(pg11) // 2
    s.write(pg11[:split])
    s.flush()
    s.flush()
    assert s.output()[-4:] == b"\x00\x00\xff\xff"
    s.write(pg11[split:split + 2])
    s.flush()
    s.write(pg11[split + 2:])
    assert zlib.decompress(s.finish(), -15) == pg11
    s = o.Stream(o.opts_default(), o.RAW)
    s.flush()
    s.write(bytes([1, 2]))
    s.flush()
    s.write(bytes([3]))
    s.flush()
    assert zlib.decompress(s.finish(), -15) == bytes([1, 2, 3])


def test_zlib_writer_checksum(pg11):
    s = o.Stream(o.opts_high(), o.ZLIB)
    s.write(pg11[:len(pg11) // 2])
    s.write(pg11[len(pg11) // 2:])
    assert s.checksum() == zlib.adler32(pg11)
    assert zlib.decompress(s.finish()) == pg11
    assert L.dfo_crc32(0, pg11, len(pg11)) == zlib.crc32(pg11)
