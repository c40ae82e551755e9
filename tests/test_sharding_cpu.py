"""CPU: host logic of the multi-rank (block-sharded) single stream, SURVEY 8(e).

The protocol code (bzip2_b200/sharding.py) is run with the oracle as its backend -- as threads for
many shapes, and once as a real world_size-2 gloo job -- and must reproduce the oracle's
single-stream output byte for byte."""
import os
import sys

import numpy as np
import pytest

import support as S
from bzip2_b200 import sharding as sh
from sharding_oracle import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split(data, world, cuts=None):
    n = data.size
    cuts = cuts or [n * (r + 1) // world for r in range(world - 1)]
    edges = [0] + list(cuts) + [n]
    return [np.ascontiguousarray(data[edges[r]:edges[r + 1]]) for r in range(world)]


@pytest.mark.parametrize("name,world,level", [
    ("text", 2, 1), ("text", 3, 1), ("mixed", 4, 1), ("runs", 2, 1), ("random", 3, 2), ("fb", 3, 1), ("small", 4, 1),
])
def test_sharded_equals_single_stream(name, world, level):
    data = {
        "text": lambda: S.gen_text(760_000),
        "mixed": lambda: S.gen_mixed(900_000, seg=90_000),
        "runs": lambda: S.gen_runs(6_000_000, seed=5),
        "random": lambda: S.gen_random(700_000),
        "fb": lambda: np.full(9_000_000, 251, np.uint8),         # one run across every shard
        "small": lambda: S.gen_text(150_000),                     # shards smaller than a block -> empty segments
    }[name]()
    shards = _split(data, world)
    halos = sh.make_halos(shards, 6_000_000)
    out, infos = sh.run_threads(world, lambda r: OracleBackend(level), shards, halos, level)
    assert out == S.orc_compress(data, level)
    segs = [i["segment"] for i in infos if i["segment"][1] > i["segment"][0]]     # non-empty segments tile the input
    assert segs[0][0] == 0 and segs[-1][1] == data.size
    for a, b in zip(segs, segs[1:]):
        assert a[1] == b[0]


def test_run_state_across_shards():
    """A shard that starts inside a long run needs the run phase of the bytes before it."""
    data = np.concatenate([S.gen_text(100_000), np.full(300_000, 7, np.uint8), S.gen_text(250_000, seed=3)])
    for cut in (100_010, 100_255, 100_256, 250_000, 399_999):
        shards = _split(data, 2, [cut])
        out, _ = sh.run_threads(2, lambda r: OracleBackend(1), shards, sh.make_halos(shards, 2_000_000), 1)
        assert out == S.orc_compress(data, 1), cut


def test_helpers():
    assert sh.fold_crcs([(1, 5)]) == 5
    a, b, c = 0x12345678, 0x9ABCDEF0, 0x0F0F0F0F
    comb = 0
    for x in (a, b, c):
        comb = (((comb << 1) | (comb >> 31)) & 0xFFFFFFFF) ^ x
    f01 = (((a << 1) | (a >> 31)) & 0xFFFFFFFF) ^ b
    assert sh.fold_crcs([(2, f01), (1, c)]) == comb
    assert sh.run_info(np.array([3, 3, 3, 4, 5, 5], np.uint8)) == (3, 3, 5, 2, 0, 6)
    assert sh.run_info(np.full(10, 9, np.uint8)) == (9, 10, 9, 10, 1, 10)
    dst = np.zeros(8, np.uint8)
    sh.or_bits(dst, 5, np.array([0xFF, 0x80], np.uint8), 9)
    assert list(dst[:2]) == [0x07, 0xFC]


def _gloo_worker(rank, world, port, path):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    import support as S2
    from bzip2_b200 import sharding as sh2
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    data = S2.gen_mixed(800_000, seg=70_000)
    shards = [np.ascontiguousarray(data[r * data.size // world:(r + 1) * data.size // world]) for r in range(world)]
    halos = sh2.make_halos(shards, 3_000_000)
    be = OracleBackend(1)
    region = np.concatenate([shards[rank], halos[rank]])
    ends = sum(s.size for s in shards[rank + 1:]) == halos[rank].size
    out, _ = sh2.compress_sharded(be, sh2.TorchComm(dist), be.load(region), int(shards[rank].size), 1, ends)
    if rank == 0:
        open(path, "wb").write(out)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    path = str(tmp_path / "sharded.bz2")
    port = 29650 + os.getpid() % 200
    mp.spawn(_gloo_worker, args=(2, port, path), nprocs=2, join=True)
    data = S.gen_mixed(800_000, seg=70_000)
    assert open(path, "rb").read() == S.orc_compress(data, 1)
