#!/usr/bin/env python3
"""Diagnostic (GPU box): latency of one dfl_compress call (host buffers) and per-stage device times by input size."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import datagen, deflate_rs_b200 as dfl, oracle_lib as o
L = dfl._native.lib()
data = datagen.silesia_mix(64 << 20)
names = (ctypes.c_char_p * 32)(); ms = (ctypes.c_float * 32)()
for size in (4 << 10, 64 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20):
    d = data[:size]
    for _ in range(3): dfl.deflate_bytes_conf(d, dfl.Compression.Default)
    t = time.perf_counter(); reps = 10
    for _ in range(reps): out = dfl.deflate_bytes_conf(d, dfl.Compression.Default)
    dt = (time.perf_counter() - t) / reps
    src = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
    dfl.compress_device(src); L.dfl_set_profiling(1); dfl.compress_device(src)
    k = L.dfl_last_stage_times(names, ms, 32); L.dfl_set_profiling(0)
    st = {names[i].decode(): round(ms[i], 3) for i in range(k)}
    t = time.perf_counter(); o.compress(d[: min(size, 4 << 20)], o.opts_default(), o.RAW); ot = (time.perf_counter() - t) * size / min(size, 4 << 20)
    print(f"{size >> 10:7d} KiB: call {dt * 1e3:8.3f} ms ({size / dt / 2**20:8.1f} MiB/s)  oracle {ot * 1e3:9.2f} ms  stages {st}")
