"""CPU checks of the GPU pipeline's algorithms (tests/model = sequential walk-through built from the
same dfl_core.h the kernels use) against the oracle.  The claim under test is DESIGN.md's central
one: sorted per-window candidate lists + per-position match records + the reference's parser run
over them + fixed 31744-token blocks reproduce the reference's stream bit for bit."""
import json
import os
import zlib

import numpy as np
import pytest

import model_lib as m
import oracle_lib as o
from conftest import FIXTURES, fixture_bytes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_symbol_arithmetic_matches_reference_tables():
    ref = json.load(open(os.path.join(GOLDEN, "ref_tables.json")))
    for stored in range(256):
        code, ne, ev, *_ = m.symbols(stored + 3, 1)
        n = ref["LENGTH_CODE"][stored]
        assert (code, ne, ev) == (257 + n, ref["LENGTH_EXTRA_BITS_LENGTH"][n], stored - ref["BASE_LENGTH"][n])
    for dist in range(1, 32769):
        *_, code, ne, ev = m.symbols(3, dist)
        want = ref["DISTANCE_CODES"][dist - 1] if dist <= 256 else ref["DISTANCE_CODES"][256 + ((dist - 1) >> 7)]
        assert (code, ne, ev) == (want, ref["DISTANCE_EXTRA_BITS"][want], dist - 1 - ref["DISTANCE_BASE"][want])


def _inputs(pg11):
    rng = np.random.default_rng(1)
    d = {
        "pg11": pg11, "short": fixture_bytes("short.bin"), "issue18": fixture_bytes("issue_18_201911.bin"),
        "dump": fixture_bytes("dump.bin"), "zeros65537": bytes(65537), "zeros61000": bytes(61000),
        "lastblock": bytes([22]) * 32768 + bytes([5, 2, 55, 11, 12]), "fives": bytes([5]) * 100000,
        "empty": b"", "one": b"\x01", "four": bytes([5, 6, 7, 8]),
        "random": rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(),
        "random4": rng.integers(0, 4, 200000, dtype=np.uint8).tobytes(),
        "period7": bytes(range(7)) * 30000,
    }
    for n in (2, 3, 5, 259, 32767, 32768, 32769, 65535, 65536, 65537, 65794, 65795):
        d[f"pg{n}"] = pg11[:n]
    for name in sorted(os.listdir(os.path.join(FIXTURES, "afl")))[:6]:
        d["afl/" + name] = fixture_bytes("afl/" + name)
    return d


@pytest.mark.parametrize("preset", list(o.PRESETS))
def test_model_is_bit_exact_with_oracle(preset, pg11):
    opts = o.PRESETS[preset]()
    for name, data in _inputs(pg11).items():
        want = o.compress(data, opts, o.RAW)
        for pseg, warm in ((8192, 1024), (512, 64)):
            got, _ = m.compress(data, opts, pseg, warm, 3)
            assert got == want, (name, preset, pseg, warm, len(got), len(want))
        assert zlib.decompress(want, -15) == data


def test_model_repairs_unsynchronised_segments():
    """Periodic data never resynchronises a speculative parse; the repair path must still be exact."""
    data = bytes(300000)
    want = o.compress(data, o.opts_default(), o.RAW)
    got, st = m.compress(data, o.opts_default(), 8192, 1024, 3)
    assert got == want and st["repairs"] + st["seq_repairs"] > 0


def test_model_custom_options(pg11):
    data = pg11[:120000]
    for checks, lazy, mt in ((4, 8, 1), (16, 258, 1), (3, 40, 1), (64, 4, 1), (7, 0, 0), (300, 64, 1), (1, 3, 1),
                             (128, 32, 1), (128, 33, 1), (129, 32, 1), (100, 20, 0)):
        opts = o.Options(checks, lazy, mt, 0)
        got, _ = m.compress(data, opts, 4096, 256, 3)
        assert got == o.compress(data, opts, o.RAW), (checks, lazy, mt)


def test_model_lazy_below_three(pg11):
    """Lazy matching with lazy_if_less_than in {0, 1, 2}: length-2 results from spurious chain entries and the
    per-call re-derivation of ignore_next matter (lz77.rs:331,374-377; matching.rs:161-165); dfl_core.h
    lz77_sequential -- the code kernel k_lz77_seq runs -- must reproduce the oracle."""
    rng = np.random.default_rng(1)
    inputs = {"pg11": pg11, "issue_18": fixture_bytes("issue_18_201911.bin"),
              "random4": rng.integers(0, 4, 200000, dtype=np.uint8).tobytes(),
              "random16": rng.integers(0, 16, 200000, dtype=np.uint8).tobytes(), "zeros": bytes(100000), "empty": b"",
              "one": b"\x01", "short": fixture_bytes("short.bin"), "pg65537": pg11[:65537]}
    for checks, lazy, mt in ((128, 0, 1), (128, 1, 1), (128, 2, 1), (1, 0, 1), (4, 2, 1), (1768, 2, 1)):
        opts = o.Options(checks, lazy, mt, 0)
        for name, data in inputs.items():
            got, _ = m.compress(data, opts, 4096, 256, 3)
            assert got == o.compress(data, opts, o.RAW), (checks, lazy, name)


def test_model_one_candidate_path_equals_the_generic_walk_and_the_oracle(pg11):
    """max_hash_checks == 1 (Compression::Fast): k_window_sort<true> settles every position against its predecessor in
    the sorted order and k_match_first the first position of every bucket against the last one of the same bucket in
    the previous window.  Same tokens as the generic walk with a budget of one, and as the oracle, across window
    boundaries (inputs of several windows), for the greedy and the lazy parser."""
    import datagen
    inputs = [pg11, datagen.silesia_mix(1 << 20), bytes(100000), datagen.enwik_like(300000)]
    for data in inputs:
        for lazy, mt in ((0, 0), (32, 1), (4, 1)):
            opts = o.Options(1, lazy, mt, 0)
            want = o.compress(data, opts, o.RAW)
            m.force_generic_match(False)
            got, _ = m.compress(data, opts)
            assert got == want, (len(data), lazy, mt, "one-candidate path")
            m.force_generic_match(True)
            try:
                gen, _ = m.compress(data, opts)
            finally:
                m.force_generic_match(False)
            assert gen == want, (len(data), lazy, mt, "generic walk")


def test_checksum_combine_arithmetic():
    """dfl_core.h crc32_combine / adler32_combine (used by the kernels' tree reductions and by the
    streaming handle) against CPython zlib on random splits, incl. empty and > 4 GiB-style lengths."""
    L = m.lib()
    rng = np.random.default_rng(11)
    for _ in range(200):
        a = rng.integers(0, 256, int(rng.integers(0, 5000)), dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, int(rng.integers(0, 70000)), dtype=np.uint8).tobytes()
        assert L.dflm_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
        assert L.dflm_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
    z = bytes(1 << 20)
    c = zlib.crc32(b"x")
    want = zlib.crc32(b"x")
    for _ in range(8):
        want = zlib.crc32(z, want)
    assert L.dflm_crc32_combine(c, zlib.crc32(z * 8), 8 << 20) == want
