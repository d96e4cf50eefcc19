#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel,
all containers, pieces and batches on inputs of a few hundred KiB."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import deflate_rs_b200 as dfl
from deflate_rs_b200 import sharding
import datagen

pg = open(os.path.join(ROOT, "tests/fixtures/pg11.txt"), "rb").read()
inputs = [pg, datagen.silesia_mix(1 << 20)[: 300000], bytes(70000), b"", b"abc", pg[:65537]]
for _ in range(1):
    for d in inputs:
        for opts in (dfl.Compression.Default, dfl.Compression.Fast, dfl.CompressionOptions.high(), dfl.CompressionOptions.rle()):
            assert zlib.decompress(dfl.deflate_bytes_conf(d, opts), -15) == d
        assert zlib.decompress(dfl.deflate_bytes_zlib(d)) == d
        assert zlib.decompress(dfl.deflate_bytes_gzip(d), 31) == d
    src = torch.frombuffer(bytearray(pg), dtype=torch.uint8).cuda()
    got = b""
    bounds = sharding.piece_bounds(len(pg), 2, 4096)
    for g, (lo, hi) in enumerate(bounds):
        out, n = sharding.encode_piece_device(src, lo, hi, dfl.Compression.Default, g == 1)
        got += bytes(out[:n].cpu().numpy())
    assert zlib.decompress(got, -15) == pg
    outs, sizes = dfl.compress_device_batch([src[:50000], src[50000:120001], src[:0]], dfl.Compression.Default, dfl.ZLIB)
    assert zlib.decompress(bytes(outs[1][:sizes[1]].cpu().numpy())) == pg[50000:120001]
    hb = dfl.compress_batch([pg[:50000], pg[50000:120001], b"", pg[:300]] * 5, dfl.Compression.Default, dfl.ZLIB)
    assert zlib.decompress(hb[5]) == pg[50000:120001] and zlib.decompress(hb[18]) == b""
    enc = dfl.write.GzEncoder(bytearray(), dfl.Compression.Default)
    enc.write_all(pg[:40000]); enc.flush(); enc.write_all(pg[40000:90000])
    assert zlib.decompress(bytes(enc.finish()), 31) == pg[:90000]
    # open pieces (no flush): parser state, uncoded tokens and the partial byte are carried over
    enc = dfl.write.ZlibEncoder(bytearray(), dfl.Compression.Default)
    enc.set_piece_bytes(30000)
    for i in range(0, len(pg), 11111):
        enc.write_all(pg[i:i + 11111])
    assert zlib.decompress(bytes(enc.finish())) == pg
    # with DFL_ONESHOT_PIECE_LIMIT set, the one-shot calls above already ran as pieces; say which it was
print("oneshot piece limit:", os.environ.get("DFL_ONESHOT_PIECE_LIMIT", "default"))
print("sanitize run ok")
