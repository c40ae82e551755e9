"""CPU: the C-ABI library loads, exports every symbol the headers declare, and fails loudly
(never silently falls back to a CPU codec) when no CUDA device is usable."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import bzip2_b200 as B
from bzip2_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(BZ2_\w+|bz2b200_\w+)\s*\(", txt))
    names.discard("bz2b200_sink")
    return names


def test_exports_match_headers(lib):
    declared = _declared("bz2_b200.h") | _declared("bzlib.h")
    assert declared, "no declarations parsed"
    assert set(binding.EXPORTS) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_oracle_in_product():
    """The product must not link, load or reference anything under oracle/."""
    out = subprocess.run(["ldd", B.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libbz2_ref" not in out
    blob = open(B.LIB_PATH, "rb").read()
    assert b"liboracle" not in blob and b"orc_compress" not in blob


def test_param_errors_need_no_gpu(lib):
    src = np.zeros(16, np.uint8)
    dst = np.zeros(700, np.uint8)
    n = C.c_uint(dst.size)
    f = lib.BZ2_bzBuffToBuffCompress
    # bzlib.c:1321-1326
    assert f(None, C.byref(n), src.ctypes.data, 16, 9, 0, 0) == binding.BZ_PARAM_ERROR
    assert f(dst.ctypes.data, None, src.ctypes.data, 16, 9, 0, 0) == binding.BZ_PARAM_ERROR
    assert f(dst.ctypes.data, C.byref(n), None, 16, 9, 0, 0) == binding.BZ_PARAM_ERROR
    assert f(dst.ctypes.data, C.byref(n), src.ctypes.data, 16, 0, 0, 0) == binding.BZ_PARAM_ERROR
    assert f(dst.ctypes.data, C.byref(n), src.ctypes.data, 16, 10, 0, 0) == binding.BZ_PARAM_ERROR
    assert f(dst.ctypes.data, C.byref(n), src.ctypes.data, 16, 9, 5, 0) == binding.BZ_PARAM_ERROR
    assert f(dst.ctypes.data, C.byref(n), src.ctypes.data, 16, 9, 0, 251) == binding.BZ_PARAM_ERROR
    strm = binding.BzStream()
    # bzlib.c:155-158
    assert lib.BZ2_bzCompressInit(None, 9, 0, 0) == binding.BZ_PARAM_ERROR
    assert lib.BZ2_bzCompressInit(C.byref(strm), 0, 0, 0) == binding.BZ_PARAM_ERROR
    assert lib.BZ2_bzCompressInit(C.byref(strm), 9, 0, 300) == binding.BZ_PARAM_ERROR
    # bzlib.c:404, :461-464
    assert lib.BZ2_bzCompress(None, 0) == binding.BZ_PARAM_ERROR
    assert lib.BZ2_bzCompress(C.byref(strm), 0) == binding.BZ_PARAM_ERROR      # state == NULL
    assert lib.BZ2_bzCompressEnd(C.byref(strm)) == binding.BZ_PARAM_ERROR


def test_fails_loudly_without_device(lib):
    if lib.bz2b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    strm = binding.BzStream()
    assert lib.BZ2_bzCompressInit(C.byref(strm), 9, 0, 0) == binding.BZ_CONFIG_ERROR
    src = np.zeros(16, np.uint8)
    dst = np.zeros(700, np.uint8)
    n = C.c_uint(dst.size)
    assert lib.BZ2_bzBuffToBuffCompress(dst.ctypes.data, C.byref(n), src.ctypes.data, 16, 9, 0, 0) == binding.BZ_CONFIG_ERROR
    with pytest.raises(B.Bz2B200Error):
        B.Engine(level=9)
    assert b"no CUDA device" in lib.bz2b200_last_error()
    # the multi-engine and input-bounded entry points fail the same way, and so does the one-shot call with a device list
    with pytest.raises(B.Bz2B200Error):
        B.Multi([0, 0], level=9)
    h = C.c_void_p()
    assert lib.bz2b200_engine_create_bounded(C.byref(h), 0, 9, 1 << 20) == -2        # BZ2B200_ENODEV
    os.environ["BZ2_B200_DEVICES"] = "0,1"
    try:
        n = C.c_uint(dst.size)
        assert lib.BZ2_bzBuffToBuffCompress(dst.ctypes.data, C.byref(n), src.ctypes.data, 16, 9, 0, 0) == binding.BZ_CONFIG_ERROR
    finally:
        del os.environ["BZ2_B200_DEVICES"]


def test_multi_param_errors_need_no_gpu(lib):
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert lib.bz2b200_multi_create(None, devs, 2, 9, 0) == -1                          # BZ2B200_EPARAM
    assert lib.bz2b200_multi_create(C.byref(h), devs, 0, 9, 0) == -1
    assert lib.bz2b200_multi_create(C.byref(h), devs, 17, 9, 0) == -1
    assert lib.bz2b200_multi_engines(None) == 0
    lib.bz2b200_multi_destroy(None)


def test_version_strings(lib):
    assert lib.BZ2_bzlibVersion().startswith(b"1.0.6")
    assert b"sm_100a" in lib.bz2b200_version()


def test_stats_struct_matches_the_header(tmp_path):
    """binding.Stats mirrors bz2b200_stats field for field: a size or offset mismatch would let the library write past it."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bz2_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(bz2b200_stats), '
                   'offsetof(bz2b200_stats, ms_total), offsetof(bz2b200_stats, out_bits), offsetof(bz2b200_stats, ms_span));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o_total, o_bits, o_span = map(int, subprocess.check_output([str(exe)]).split())
    assert C.sizeof(binding.Stats) == size
    assert binding.Stats.ms_total.offset == o_total and binding.Stats.out_bits.offset == o_bits and binding.Stats.ms_span.offset == o_span
