"""CPU-only checks of the boundary: the shared library loads, exports every symbol the public header
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import deflate_rs_b200 as dfl
from deflate_rs_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    return os.path.exists(_native.LIB_PATH)


@pytest.fixture(scope="module", autouse=True)
def _build():
    if not _built():
        import __graft_entry__ as g
        g.build()


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "deflate_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dfl_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    L = _native.lib()
    for name in sorted(declared):
        assert getattr(L, name) is not None, name


def test_presets_match_reference_constants():
    """compression_options.rs:14-20,126-178"""
    L = _native.lib()
    want = {0: (1, 0, 0), 1: (128, 32, 1), 2: (1768, 128, 1), 3: (0, 0, 0), 4: (0, 0, 1)}
    for preset, (checks, lazy, mt) in want.items():
        o = _native.dfl_options()
        assert L.dfl_options_preset(preset, ctypes.byref(o)) == 0
        assert (o.max_hash_checks, o.lazy_if_less_than, o.matching_type, o.special) == (checks, lazy, mt, 0)
    assert L.dfl_options_preset(99, ctypes.byref(_native.dfl_options())) < 0
    assert dfl.CompressionOptions.from_(dfl.Compression.Best) == dfl.CompressionOptions.high()
    assert dfl.CompressionOptions.from_(dfl.Compression.Default) == dfl.CompressionOptions.default()
    assert dfl.CompressionOptions.from_(dfl.Compression.Fast) == dfl.CompressionOptions.fast()
    assert dfl.CompressionOptions.rle().max_hash_checks == 0 and dfl.CompressionOptions.rle().matching_type == dfl.MatchingType.Lazy


def test_bound_and_strerror():
    L = _native.lib()
    assert L.dfl_version() == 100
    for n in (0, 1, 31744, 32767, 1 << 20, 1 << 30):
        assert L.dfl_bound(n, _native.RAW) >= n + 5 * (n // 32767 + 1) + 2
    assert b"no CPU fallback" in L.dfl_strerror(-4)
    assert L.dfl_strerror(0) == b"ok"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert _native.lib().dfl_device_count() == 0
    with pytest.raises(dfl.DeflateB200Error) as ei:
        dfl.deflate_bytes(b"hello hello hello")
    assert ei.value.status == -4
    enc = dfl.write.DeflateEncoder(bytearray(), dfl.Compression.Default)
    assert enc.write(b"abc") == 3          # buffering input is host logic
    with pytest.raises(dfl.DeflateB200Error):
        enc.finish()                        # compressing needs the device


def test_argument_validation():
    L = _native.lib()
    n = ctypes.c_size_t()
    o = _native.dfl_options(128, 32, 1, 0)
    assert L.dfl_compress(None, 5, ctypes.byref(o), 0, None, 0, None, 0, ctypes.byref(n)) == -1
    assert L.dfl_encoder_new(ctypes.byref(o), 7, None, 0) is None
    with pytest.raises(TypeError):
        dfl.deflate_bytes_conf(b"x", "fast")
    with pytest.raises(ValueError):
        dfl.CompressionOptions(max_hash_checks=70000)._c()


def _build_c_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "deflate_file")
    lib_dir = os.path.dirname(_native.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "deflate_file.c"), "-o", exe, "-L", lib_dir, "-ldeflate_b200",
                           "-Wl,-rpath," + lib_dir])
    return exe


def test_c_caller_links_against_the_header_and_fails_loudly_without_a_device(tmp_path):
    """examples/deflate_file.c is a plain-C caller of include/deflate_b200.h: it must compile against the header
    as shipped (no C++-isms), link with the in-tree library, and -- on a box without a GPU -- stop with the
    library's "no CUDA device" message instead of producing output."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU variant")
    exe = _build_c_example(tmp_path)
    out = tmp_path / "out.z"
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "fixtures", "pg11.txt"), str(out)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
    assert not out.exists()


@pytest.mark.gpu
def test_c_caller_output_equals_the_oracle(tmp_path):
    import subprocess
    import zlib
    import oracle_lib as o
    exe = _build_c_example(tmp_path)
    src = os.path.join(ROOT, "tests", "fixtures", "pg11.txt")
    data = open(src, "rb").read()
    for level, opts in (("default", o.opts_default()), ("fast", o.opts_fast())):
        for wrap, owrap, wbits in (("raw", o.RAW, -15), ("zlib", o.ZLIB, 15), ("gzip", o.GZIP, 31)):
            out = tmp_path / f"{level}.{wrap}"
            subprocess.check_call([exe, src, str(out), level, wrap])
            got = out.read_bytes()
            assert zlib.decompress(got, wbits) == data
            assert got == o.compress(data, opts, owrap), (level, wrap)


def _build_cpp_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "deflate_file_cpp")
    lib_dir = os.path.dirname(_native.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "deflate_file.cpp"), "-o", exe, "-L", lib_dir, "-ldeflate_b200",
                           "-Wl,-rpath," + lib_dir])
    return exe


def test_cpp_header_compiles_and_fails_loudly_without_a_device(tmp_path):
    """include/deflate_b200.hpp (the C++ mirror of the crate API, SURVEY 8(b) "who calls it") must compile with
    -Wall -Werror against the C header as shipped, and on a box without a GPU stop with the library's message."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU variant")
    exe = _build_cpp_example(tmp_path)
    out = tmp_path / "out.zlib"
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "fixtures", "pg11.txt"), str(out)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
    assert not out.exists()


@pytest.mark.gpu
def test_cpp_caller_output_equals_the_oracle(tmp_path):
    import subprocess
    import oracle_lib as o
    exe = _build_cpp_example(tmp_path)
    src = os.path.join(ROOT, "tests", "fixtures", "pg11.txt")
    out = tmp_path / "out.zlib"
    subprocess.check_call([exe, src, str(out)])   # the example itself checks one-shot == streamed
    assert out.read_bytes() == o.compress(open(src, "rb").read(), o.opts_default(), o.ZLIB)


def test_rust_shim_binds_only_symbols_the_library_exports():
    """rust-shim/src/lib.rs cannot be compiled here (no rustc); at least every `pub fn dfl_*` it declares must be
    a symbol include/deflate_b200.h declares and the built library exports."""
    import re
    src = open(os.path.join(ROOT, "rust-shim", "src", "lib.rs")).read()
    names = set(re.findall(r"pub fn (dfl_[a-z0-9_]+)\(", src))
    assert len(names) >= 20
    hdr = open(os.path.join(ROOT, "include", "deflate_b200.h")).read()
    L = _native.lib()
    for nm in sorted(names):
        assert re.search(r"\b" + nm + r"\(", hdr), nm
        assert hasattr(L, nm), nm
