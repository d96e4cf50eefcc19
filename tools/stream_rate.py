"""Throughput of the streaming handle (writer.rs path): dfl_encoder_write in fixed-size writes + finish,
host memory in and out, against the one-shot host call on the same input.
Usage (GPU box): python tools/stream_rate.py [size_mib] [write_kib] [preset]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
import deflate_rs_b200 as dfl  # noqa: E402

size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
wr = (int(sys.argv[2]) if len(sys.argv) > 2 else 4096) << 10
preset = sys.argv[3] if len(sys.argv) > 3 else "default"
opts = {"default": dfl.CompressionOptions.default(), "fast": dfl.CompressionOptions.fast(),
        "high": dfl.CompressionOptions.high()}[preset]._c()
L = dfl._native.lib()
data = datagen.silesia_mix(size, 0x51DE51A)
buf = (ctypes.c_uint8 * size).from_buffer_copy(data)
base = ctypes.addressof(buf)
cap = L.dfl_bound(size, dfl.ZLIB) + 64
out = (ctypes.c_uint8 * cap)()
n_out = ctypes.c_size_t()


def oneshot():
    rc = L.dfl_compress(base, size, ctypes.byref(opts), dfl.ZLIB, None, 0, out, cap, ctypes.byref(n_out))
    assert rc == 0
    return bytes(memoryview(out)[:n_out.value])


def stream():
    e = L.dfl_encoder_new(ctypes.byref(opts), dfl.ZLIB, None, 0)
    got = 0
    p = ctypes.POINTER(ctypes.c_uint8)()
    ln = ctypes.c_size_t()
    chunks = []
    for off in range(0, size, wr):
        rc = L.dfl_encoder_write(e, base + off, min(wr, size - off), None)
        assert rc == 0
        L.dfl_encoder_take_output(e, ctypes.byref(p), ctypes.byref(ln))
        if ln.value:
            chunks.append(ctypes.string_at(p, ln.value))
            L.dfl_encoder_advance_output(e, ln.value)
    assert L.dfl_encoder_flush(e, dfl._native.FLUSH_FINISH) == 0
    L.dfl_encoder_take_output(e, ctypes.byref(p), ctypes.byref(ln))
    chunks.append(ctypes.string_at(p, ln.value))
    L.dfl_encoder_free(e)
    return b"".join(chunks)


for name, fn in (("oneshot", oneshot), ("stream", stream)):
    ref = fn()
    fn()
    t0 = time.perf_counter()
    r = fn()
    dt = time.perf_counter() - t0
    print(f"{name}: {size / dt / 2**20:.0f} MiB/s ({dt * 1e3:.1f} ms), {len(r)} bytes")
    if name == "oneshot":
        first = ref
    else:
        assert ref == first, "stream and one-shot differ"
print("identical")
