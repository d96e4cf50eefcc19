#!/usr/bin/env python3
"""Diagnostic (GPU box): stage times per kind of synthetic data (64 MiB of each)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import datagen, deflate_rs_b200 as dfl
L = dfl._native.lib()
size = 64 << 20
rng = np.random.default_rng(5)
kinds = {
    "text": lambda: b"".join(datagen._zipf_text(rng, 1 << 20) for _ in range(size >> 20)),
    "xml": lambda: b"".join(datagen._xml_records(rng, 1 << 20) for _ in range(size >> 20)),
    "binary": lambda: b"".join(datagen._binary_records(rng, 1 << 20) for _ in range(size >> 20)),
    "sparse": lambda: b"".join(datagen._sparse(rng, 1 << 20) for _ in range(size >> 20)),
    "random": lambda: b"".join(datagen._random(rng, 1 << 20) for _ in range(size >> 20)),
    "mix": lambda: datagen.silesia_mix(size),
}
names = (ctypes.c_char_p * 32)(); ms = (ctypes.c_float * 32)()
for kind, gen in kinds.items():
    data = gen()
    src = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    row = [kind]
    for _ in range(2):
        dfl.compress_device(src, dfl.Compression.Default, dfl.RAW)
    L.dfl_set_profiling(1)
    out, n = dfl.compress_device(src, dfl.Compression.Default, dfl.RAW)
    k = L.dfl_last_stage_times(names, ms, 32)
    L.dfl_set_profiling(0)
    st = {names[i].decode(): ms[i] for i in range(k)}
    row.append(" ".join(f"{a} {b:.2f}" for a, b in st.items()) + f" | total {sum(st.values()):.2f} ms ratio {n/len(data):.3f}")
    print(" | ".join(row))
