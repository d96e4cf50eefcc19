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


def stream_phases():
    """the same stream, timed by phase; output taken once at the end"""
    e = L.dfl_encoder_new(ctypes.byref(opts), dfl.ZLIB, None, 0)
    p = ctypes.POINTER(ctypes.c_uint8)()
    ln = ctypes.c_size_t()
    t0 = time.perf_counter()
    for off in range(0, size, wr):
        assert L.dfl_encoder_write(e, base + off, min(wr, size - off), None) == 0
    t1 = time.perf_counter()
    assert L.dfl_encoder_flush(e, dfl._native.FLUSH_FINISH) == 0
    t2 = time.perf_counter()
    L.dfl_encoder_take_output(e, ctypes.byref(p), ctypes.byref(ln))
    r = ctypes.string_at(p, ln.value)
    t3 = time.perf_counter()
    L.dfl_encoder_free(e)
    print(f"  phases: writes {1e3 * (t1 - t0):.1f} ms, finish {1e3 * (t2 - t1):.1f} ms, copy out of the handle {1e3 * (t3 - t2):.1f} ms"
          f" -> {size / (t2 - t0) / 2**20:.0f} MiB/s in the library")
    return r


def stream_drain():
    """the stream with its output handed over after every write (as a Write sink sees it) but not copied again"""
    e = L.dfl_encoder_new(ctypes.byref(opts), dfl.ZLIB, None, 0)
    p = ctypes.POINTER(ctypes.c_uint8)()
    ln = ctypes.c_size_t()
    total = 0
    t0 = time.perf_counter()
    for off in range(0, size, wr):
        assert L.dfl_encoder_write(e, base + off, min(wr, size - off), None) == 0
        L.dfl_encoder_take_output(e, ctypes.byref(p), ctypes.byref(ln))
        total += ln.value
        L.dfl_encoder_advance_output(e, ln.value)
    assert L.dfl_encoder_flush(e, dfl._native.FLUSH_FINISH) == 0
    L.dfl_encoder_take_output(e, ctypes.byref(p), ctypes.byref(ln))
    total += ln.value
    t1 = time.perf_counter()
    L.dfl_encoder_free(e)
    print(f"  drained, not copied: {size / (t1 - t0) / 2**20:.0f} MiB/s ({1e3 * (t1 - t0):.1f} ms), {total} bytes")


def raw_copies():
    import torch
    h = torch.frombuffer(buf, dtype=torch.uint8)
    d = torch.empty(size, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h); torch.cuda.synchronize(); t1 = time.perf_counter()
        h2 = d[:size // 3].cpu(); t2 = time.perf_counter()
    print(f"  pageable copies on this box: H2D {size / (t1 - t0) / 2**30:.1f} GiB/s, D2H {size / 3 / (t2 - t1) / 2**30:.1f} GiB/s")


raw_copies()
stream_phases()
stream_phases()
stream_drain()
stream_drain()
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
