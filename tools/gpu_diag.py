#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA path with the oracle (run on a GPU box; prints a report).
Used while bringing kernels up: it localises a mismatch to the LZ77 stage (token stream) or to the
entropy stage (block coding / bit packing) and prints the first difference."""
import ctypes
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import deflate_rs_b200 as dfl  # noqa: E402
import oracle_lib as o  # noqa: E402

L = dfl._native.lib()


def gpu_tokens(data, opts):
    got = np.zeros(len(data) + 16, dtype=np.uint32)
    n = ctypes.c_size_t()
    co = dfl._native.dfl_options(opts.max_hash_checks, opts.lazy_if_less_than, opts.matching_type, 0)
    rc = L.dfl_lz77_tokens(data, len(data), ctypes.byref(co), got.ctypes.data_as(ctypes.c_void_p), len(got), ctypes.byref(n))
    return rc, got[: n.value]


def oracle_tokens(data, opts):
    litlen, dist, ends = o.lz77_tokens(data, opts)
    return np.where(dist == 0, litlen, (litlen + 3) | (dist << 9)).astype(np.uint32)


def gpu_entropy(data, toks):
    cap = L.dfl_bound(len(data), 0)
    out = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t()
    rc = L.dfl_encode_tokens(data, len(data), toks.ctypes.data_as(ctypes.c_void_p), len(toks), out, cap, ctypes.byref(n))
    return rc, out.raw[: n.value]


def first_diff(a, b):
    m = min(len(a), len(b))
    for i in range(m):
        if a[i] != b[i]:
            return i
    return m if len(a) != len(b) else -1


def report(name, data, preset):
    opts = o.PRESETS[preset]()
    want = o.compress(data, opts, o.RAW)
    wt = oracle_tokens(data, opts)
    rc, gt = gpu_tokens(data, opts)
    cnt = (ctypes.c_uint64 * 8)()
    L.dfl_last_counters(cnt, 8)
    tok_ok = rc == 0 and len(gt) == len(wt) and bool((gt == wt).all())
    line = f"{name:14s} {preset:12s} n={len(data):9d} tokens rc={rc} {'OK ' if tok_ok else 'BAD'} ({len(gt)} vs {len(wt)}) repairs par={cnt[3]} seq={cnt[4]}"
    if not tok_ok and rc == 0:
        i = first_diff(gt.tolist(), wt.tolist())
        pos = int(sum((t & 0x1ff) if (t >> 9) else 1 for t in wt[:i].tolist()))
        line += f" first token diff at #{i} (input pos {pos}): gpu={gt[i] if i < len(gt) else None:#x} oracle={wt[i] if i < len(wt) else None:#x}"
    print(line)
    rc, ge = gpu_entropy(data, wt)
    ent_ok = rc == 0 and ge == want
    line = f"{'':14s} {'':12s} entropy rc={rc} {'OK ' if ent_ok else 'BAD'} ({len(ge)} vs {len(want)})"
    if not ent_ok and rc == 0:
        i = first_diff(ge, want)
        line += f" first byte diff at {i}"
        try:
            line += f" inflate_ok={zlib.decompress(ge, -15) == data}"
        except Exception as e:
            line += f" inflate_error={e}"
    print(line)
    try:
        t = time.time()
        full = dfl.deflate_bytes_conf(data, dfl.CompressionOptions(opts.max_hash_checks, opts.lazy_if_less_than, dfl.MatchingType(opts.matching_type)))
        dt = time.time() - t
        ok = full == want
        rt = zlib.decompress(full, -15) == data
        print(f"{'':14s} {'':12s} full   {'OK ' if ok else 'BAD'} roundtrip={rt} ({len(full)} vs {len(want)}) {dt * 1e3:.1f} ms")
    except Exception as e:
        print(f"{'':14s} {'':12s} full   EXC {e}")
    sys.stdout.flush()


def main():
    fx = os.path.join(ROOT, "tests", "fixtures")
    pg11 = open(os.path.join(fx, "pg11.txt"), "rb").read()
    rng = np.random.default_rng(1)
    inputs = {
        "six": bytes([10, 10, 10, 10, 10, 55]), "empty": b"", "gnu": b"                    GNU GENERAL PUBLIC LICENSE",
        "pg11[:3000]": pg11[:3000], "pg11[:40000]": pg11[:40000], "pg11": pg11,
        "zeros100k": bytes(100000), "random150k": rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(),
    }
    for name, data in inputs.items():
        for preset in ("fast", "default", "high", "rle", "huffman_only"):
            report(name, data, preset)


if __name__ == "__main__":
    main()
