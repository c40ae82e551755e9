"""GPU: several engines on one stream (bz2b200_multi_*, csrc/multi.cu) give the single-engine bytes.

On a one-GPU box the engines share the GPU (devices [0, 0, ...]): the chains, the prefetch guesses, the seam bytes and the
trailer are exercised exactly as across GPUs; with more GPUs visible the same tests spread over them."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import support as S
import bzip2_b200 as B

pytestmark = pytest.mark.gpu
G = S.GOLDEN


def _devices(k):
    n = max(1, B.load().bz2b200_device_count())
    return [i % n for i in range(k)]


@pytest.fixture(scope="module")
def small_multi():
    """three engines with 8 MiB windows at -1: tens of windows per call"""
    ms = {}

    def get(level, k=3, window=8 << 20):
        key = (level, k, window)
        if key not in ms:
            for m in ms.values():
                m.close()
            ms.clear()
            ms[key] = B.Multi(_devices(k), level=level, window_bytes=window)
        return ms[key]
    yield get
    for m in ms.values():
        m.close()


def test_multi_equals_single_many_windows(engine_for, small_multi):
    m = small_multi(1)
    for name, d in [("mixed", S.gen_mixed(70_000_000, seg=1 << 22)), ("text", S.gen_text(33_000_000, seed=3)),
                    ("runs", S.gen_runs(120_000_000, seed=4)), ("fb", np.full(50_000_001, 251, np.uint8)),
                    ("tiny", S.gen_text(1000)), ("one", np.array([7], np.uint8)), ("empty", np.zeros(0, np.uint8))]:
        exp = engine_for(1).compress(d)
        got = m.compress(d)
        assert got == exp, name
        if d.size > 30_000_000:
            assert m.stats.n_windows >= 3, name
            assert m.stats.n_blocks == engine_for(1).stats.n_blocks


def test_multi_tail_streamed_and_outfull(engine_for, small_multi):
    m = small_multi(1)
    nmax = 99981
    d = (np.arange(nmax + 1, dtype=np.uint32) % 251).astype(np.uint8)
    assert m.compress(d) == S.orc_compress(d, 1)
    assert m.compress(d, flags=1) == S.orc_compress(d, 1, tail_merge=0)
    big = S.gen_random(20_000_000)
    out = np.empty(1_000_000, np.uint8)
    with pytest.raises(B.Bz2B200Error):
        m.compress_ptr(big.ctypes.data, big.size, out.ctypes.data, out.size)
    # the engines stay usable after a failed call
    assert m.compress(d) == S.orc_compress(d, 1)


def test_multi_level9_golden(small_multi):
    """Default windows, two and four engines: the 200 MB C4 golden of the reference (tests/golden/large_streams.json)."""
    gold = json.load(open(os.path.join(G, "large_streams.json")))["c4_200M_L9"]
    d = S.gen_c4(200_000_000, seg=64 << 20)
    for k in (2, 4):
        m = small_multi(9, k=k, window=32 << 20 if k == 4 else 0)
        out = m.compress(d)
        assert len(out) == gold["out_len"] and hashlib.sha256(out).hexdigest() == gold["sha256"], k


def test_multi_resident_input(engine_for, small_multi):
    torch = pytest.importorskip("torch")
    m = small_multi(1)
    d = S.gen_mixed(40_000_000, seg=1 << 21)
    exp = engine_for(1).compress(d)
    copies = {}
    ptrs = []
    for dev in m.devices:
        if dev not in copies:
            copies[dev] = torch.from_numpy(d).to(f"cuda:{dev}")
        ptrs.append(copies[dev].data_ptr())
    cap = m.out_cap(d.size)
    out = np.empty(cap, np.uint8)
    n = m.compress_ptr(None, d.size, out.ctypes.data, cap, d_srcs=ptrs)
    assert out[:n].tobytes() == exp


def test_buff_to_buff_uses_the_device_list(engine_for, monkeypatch):
    """BZ2_B200_DEVICES with several entries routes BZ2_bzBuffToBuffCompress through the engines of that list."""
    lib = B.load()
    d = S.gen_mixed(60_000_000, seg=1 << 22)
    exp = engine_for(2).compress(d)
    monkeypatch.setenv("BZ2_B200_DEVICES", ",".join(str(x) for x in _devices(2)))
    monkeypatch.setenv("BZ2_B200_WINDOW_MB", "16")
    lib.bz2b200_pool_clear()
    try:
        assert B.compress(d, 2) == exp
        assert B.compress(d[:5_000_000], 2) == engine_for(2).compress(d[:5_000_000])
    finally:
        monkeypatch.delenv("BZ2_B200_DEVICES")
        monkeypatch.delenv("BZ2_B200_WINDOW_MB")
        lib.bz2b200_pool_clear()


def test_plain_c_client_with_device_list(engine_for, tmp_path):
    """A C program that only knows bzlib.h: BZ2_B200_DEVICES in its environment spreads the one-shot call over the engines."""
    import subprocess
    exe = os.path.join(S.ROOT, "tests", "c", "multi_buff")
    if not os.path.exists(exe):
        pytest.skip("tests/c/multi_buff not built")
    d = S.gen_c4(150_000_000, seg=16 << 20)
    src = tmp_path / "in.dat"
    src.write_bytes(d.tobytes())
    env = dict(os.environ, BZ2_B200_DEVICES=",".join(str(x) for x in _devices(4)), BZ2_B200_WINDOW_MB="24")
    r = subprocess.run([exe, str(src), "9", "2"], capture_output=True, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout == engine_for(9).compress(d)
