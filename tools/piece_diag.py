#!/usr/bin/env python3
"""Diagnostic (GPU box): encode pieces of one stream independently and compare every piece with the
oracle writer's output between flushes; prints the first differing piece."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import deflate_rs_b200 as dfl
from deflate_rs_b200 import sharding
import oracle_lib as o

data = open(os.path.join(ROOT, "tests/fixtures/pg11.txt"), "rb").read()[:int(sys.argv[1]) if len(sys.argv) > 1 else 50000]
piece = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
bounds = [(lo, min(len(data), lo + piece)) for lo in range(0, len(data), piece)]
src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
s = o.Stream(o.opts_default(), o.RAW)
seen = 0
for g, (lo, hi) in enumerate(bounds):
    last = g + 1 == len(bounds)
    out, n = sharding.encode_piece_device(src, lo, hi, dfl.Compression.Default, last)
    got = bytes(out[:n].cpu().numpy())
    s.write(data[lo:hi])
    if last: full = s.finish()
    else:
        s.flush(); full = s.output()
    want = full[seen:]; seen = len(full)
    same = got == want
    print(f"piece {g} [{lo},{hi}) got {len(got)} want {len(want)} {'OK' if same else 'DIFF'}")
    if not same:
        # decode both with the dictionary to see where the token streams part
        for name, blob in (("got", got), ("want", want)):
            d = zlib.decompressobj(-15, zdict=data[max(0, lo - 32768):lo]) if lo else zlib.decompressobj(-15)
            try:
                dec = d.decompress(blob)
                print("  ", name, "inflates to piece:", dec == data[lo:hi], len(dec))
            except Exception as e:
                print("  ", name, "inflate error", e)
        k = next(i for i in range(min(len(got), len(want))) if got[i] != want[i]) if got[:min(len(got), len(want))] != want[:min(len(got), len(want))] else min(len(got), len(want))
        print("   first diff at byte", k)
        break
