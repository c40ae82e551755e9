"""One .bz2 stream sharded by block over several ranks (one process per GPU), SURVEY.md 8(e).

Blocks are independent once their boundaries are known; boundaries are a greedy chain over the
RLE1 chunk structure.  Protocol (no data-path collective, only scalars and the finished pieces move):

  1. every rank scans its shard plus a halo of the following data for chunk ends     (parallel)
  2. rank r receives the input offset of its first block from rank r-1, walks the chain to the
     first boundary at or past its shard end, and hands that offset to rank r+1      (one integer)
  3. every rank compresses its block-aligned segment as a header-less, trailer-less piece (parallel)
  4. bit lengths / block counts / CRC folds are all-gathered; rank 0 receives the pieces and
     bit-shifts them into place behind the stream header, then appends the trailer     (S5)

A backend supplies load / scan / boundary / compress_segment / new_stream / place / to_host.  The product backend
is GpuBackend (libbz2_b200.so); the CPU tests of this host logic (gloo, world_size 2) plug in a checker backend that
lives with the tests (tests/sharding_oracle.py) -- nothing in this package imports the oracle.

For a single process that owns several GPUs the same sharding is available behind the C API without torch:
bz2b200_multi_* (csrc/multi.cu), which BZ2_bzBuffToBuffCompress uses when BZ2_B200_DEVICES lists several devices.
"""
import ctypes as C

import numpy as np

NO_HEADER, NO_TRAILER, TAIL_STREAMED = 2, 4, 1


def run_info(a):
    """(first byte, leading run, last byte, trailing run, all-same, length) of a byte array."""
    n = int(a.size)
    if n == 0:
        return (256, 0, 256, 0, 1, 0)

    def run_from(arr, value):
        w = 4096
        done = 0
        while done < arr.size:
            chunk = arr[done:done + w]
            bad = np.nonzero(chunk != value)[0]
            if bad.size:
                return done + int(bad[0])
            done += chunk.size
            w *= 8
        return int(arr.size)

    lead = run_from(a, a[0])
    trail = run_from(a[::-1], a[-1])
    return (int(a[0]), lead, int(a[-1]), trail, 1 if lead == n else 0, n)


def prev_state(infos, r):
    """Byte before shard r and the length of the run it ends, from the run_info of shards 0..r-1."""
    q = r - 1
    while q >= 0 and infos[q][5] == 0:
        q -= 1
    if q < 0:
        return 256, 0
    byte, run = infos[q][2], infos[q][3]
    while infos[q][4] and q > 0:
        q -= 1
        if infos[q][5] == 0:
            continue
        if infos[q][2] != byte:
            break
        run += infos[q][3]
        if not infos[q][4]:
            break
    return byte, run


def fold_crcs(parts):
    """combinedCRC over segments: parts = [(n_blocks, fold_from_zero), ...] (compress.c:826-828)."""
    comb = 0
    for nblk, fold in parts:
        k = nblk % 32
        comb = ((comb << k) | (comb >> (32 - k))) & 0xFFFFFFFF if k else comb
        comb ^= fold
    return comb


def or_bits(dst, dst_bit, src, nbits):
    """numpy S5 (used by the CPU backend and as the checker of bz2b200_concat_bits)."""
    if nbits == 0:
        return
    bits = np.unpackbits(src[: (nbits + 7) // 8])[:nbits]
    lo = dst_bit // 8
    hi = (dst_bit + nbits + 7) // 8
    window = np.unpackbits(dst[lo:hi])
    off = dst_bit - lo * 8
    window[off:off + nbits] |= bits
    dst[lo:hi] = np.packbits(window)


# ------------------------------------------------------------------------------------ backends
class GpuBackend:
    """libbz2_b200.so on this rank's GPU; arrays are torch uint8 tensors on the device."""

    def __init__(self, level, device):
        import torch
        from . import binding
        self.torch, self.binding, self.level, self.device = torch, binding, level, device
        self.lib = binding.load()
        self.eng = binding.Engine(level=level, device=device)
        self.dev = torch.device("cuda", device)
        self.out = None

    def load(self, region):
        t = self.torch.from_numpy(np.ascontiguousarray(region, np.uint8))
        return t.to(self.dev, non_blocking=False)

    def scan(self, region, prev_byte, prev_run, input_ends):
        # the scan's buffers are kept between calls (a step loop pays no cudaMalloc / cudaFree)
        n = region.numel()
        h = getattr(self, "_scan", None)
        if h is not None and n <= self._scan_cap:
            rc = self.lib.bz2b200_scan_rescan(h, region.data_ptr(), n, prev_byte, prev_run, 1 if input_ends else 0)
            if rc:
                raise self.binding.Bz2B200Error(f"scan_rescan rc={rc}: {self.lib.bz2b200_last_error().decode()}")
            return h
        if h is not None:
            self.lib.bz2b200_scan_destroy(h)
            self._scan = None
        h = C.c_void_p()
        rc = self.lib.bz2b200_scan_create(C.byref(h), self.device, self.level, region.data_ptr(), n,
                                          prev_byte, prev_run, 1 if input_ends else 0)
        if rc:
            raise self.binding.Bz2B200Error(f"scan_create rc={rc}: {self.lib.bz2b200_last_error().decode()}")
        self._scan, self._scan_cap = h, n
        return h

    def boundary(self, scan, start, limit, tail_streamed):
        b, nb = C.c_size_t(0), C.c_uint32(0)
        rc = self.lib.bz2b200_scan_boundary(scan, start, limit, TAIL_STREAMED if tail_streamed else 0, C.byref(b), C.byref(nb))
        if rc:
            raise self.binding.Bz2B200Error(f"scan_boundary rc={rc}: {self.lib.bz2b200_last_error().decode()}")
        return int(b.value), int(nb.value)

    def free_scan(self, scan):
        pass                                   # kept for the next call; released by close()

    def close(self):
        if getattr(self, "_scan", None) is not None:
            self.lib.bz2b200_scan_destroy(self._scan)
            self._scan = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def compress_segment(self, region, start, end, flags):
        n = end - start
        cap = n + n // 50 + 24576 * (n // (100000 * self.level - 19) + 2) + 1024
        cap = (cap + 255) & ~255
        if self.out is None or self.out.numel() < cap:
            self.out = self.torch.empty(cap, dtype=self.torch.uint8, device=self.dev)
        if n == 0:
            self.out[:64].zero_()
            return self.out, 0, 0, 0
        self.eng.compress_device(region.data_ptr() + start, n, self.out.data_ptr(), self.out.numel(), flags=flags)
        st = self.eng.stats
        return self.out, int(st.out_bits), int(st.n_blocks), int(st.combined_crc)

    def new_stream(self, nbytes):
        return self.torch.zeros((nbytes + 8 + 3) & ~3, dtype=self.torch.uint8, device=self.dev)

    def place(self, stream, bit, piece, nbits):
        if nbits == 0:
            return
        if not hasattr(piece, "data_ptr"):
            arr = np.zeros((len(piece) + 7) & ~3, np.uint8)
            arr[: len(piece)] = np.frombuffer(bytes(piece), np.uint8)
            piece = self.torch.from_numpy(arr).to(self.dev)
        rc = self.lib.bz2b200_concat_bits(self.device, stream.data_ptr(), bit, piece.data_ptr(), nbits)
        if rc:
            raise self.binding.Bz2B200Error(f"concat_bits rc={rc}: {self.lib.bz2b200_last_error().decode()}")

    def to_host(self, stream, nbytes):
        # pinned staging buffer, reused between calls; the result is a read-only view of it
        if getattr(self, "_pin", None) is None or self._pin.numel() < nbytes:
            self._pin = self.torch.empty(nbytes + (nbytes >> 4) + 4096, dtype=self.torch.uint8).pin_memory()
        self._pin[:nbytes].copy_(stream[:nbytes], non_blocking=True)
        self.torch.cuda.current_stream(self.dev).synchronize()
        return memoryview(self._pin.numpy())[:nbytes]


# ------------------------------------------------------------------------------------ protocol
def compress_sharded(backend, comm, region, shard_len, level, stream_ends_in_region, tail_streamed=False, return_host=True):
    """Compress this rank's shard of the stream.

    region: shard followed by its halo (already in the backend's memory: backend.load()).
    shard_len: bytes of `region` that belong to this rank.
    stream_ends_in_region: the region reaches the end of the whole stream.
    comm: rank, world, all_gather_ints(list[int]) -> list[list[int]], send_int(dst, v), recv_int(src) -> int,
          send_piece(dst, piece, nbytes), recv_piece(src, nbytes) -> piece.
    Returns the finished stream (bytes) on rank 0 (None elsewhere) and a dict of counters.
    """
    import time as _time
    rank, world = comm.rank, comm.world
    marks = [("start", _time.perf_counter())]
    mark = lambda name: marks.append((name, _time.perf_counter()))
    n_region = int(region.numel() if hasattr(region, "numel") else region.size)
    host_view = comm.host_view(region, shard_len)
    infos = comm.all_gather_ints(list(run_info(host_view)))
    mark("run_info")
    prev_byte, prev_run = prev_state(infos, rank)
    lens = [i[5] for i in infos]
    base = sum(lens[:rank])                              # global offset of this shard
    scan = backend.scan(region, prev_byte, prev_run, stream_ends_in_region) if shard_len else None
    mark("scan")
    # 2. the chain: global offset of my first block boundary
    start_g = 0 if rank == 0 else comm.recv_int(rank - 1)
    last = rank == world - 1
    if shard_len == 0 or start_g >= base + shard_len:
        seg = (0, 0)
        next_g = start_g
    else:
        s_loc = start_g - base
        limit = n_region if last else shard_len
        b_loc, _ = backend.boundary(scan, s_loc, limit, tail_streamed)
        seg = (s_loc, b_loc)
        next_g = base + b_loc
    if not last:
        comm.send_int(rank + 1, next_g)
    mark("chain")
    if scan is not None and hasattr(backend, "free_scan"):
        backend.free_scan(scan)
    # 3. my piece
    flags = NO_HEADER | NO_TRAILER | (TAIL_STREAMED if tail_streamed else 0)
    piece, nbits, nblk, fold = backend.compress_segment(region, seg[0], seg[1], flags)
    mark("compress")
    # 4. assembly on rank 0
    meta = comm.all_gather_ints([nbits, nblk, fold])
    mark("meta")
    offs = [32]
    for m in meta:
        offs.append(offs[-1] + m[0])
    total_bits = offs[-1] + 80
    nbytes = (total_bits + 7) // 8
    info = {"segment": (base + seg[0], base + seg[1]), "nbits": nbits, "blocks": nblk, "total_bytes": nbytes}

    def finish_marks():
        info["protocol_ms"] = {b[0]: round((b[1] - a[1]) * 1e3, 3) for a, b in zip(marks, marks[1:])}

    if rank != 0:
        comm.send_piece(0, piece, (nbits + 7) // 8)
        mark("send_piece")
        finish_marks()
        return None, info
    stream = backend.new_stream(nbytes)
    header = bytes([0x42, 0x5A, 0x68, 0x30 + level])
    backend.place(stream, 0, header, 32)
    backend.place(stream, offs[0], piece, nbits)
    for r in range(1, world):
        pr = comm.recv_piece(r, (meta[r][0] + 7) // 8)
        backend.place(stream, offs[r], pr, meta[r][0])
    comb = fold_crcs([(m[1], m[2]) for m in meta])
    trailer = (0x177245385090 << 32) | comb
    backend.place(stream, offs[-1], trailer.to_bytes(10, "big"), 80)
    info["combined_crc"] = comb
    mark("assemble")
    out = backend.to_host(stream, nbytes) if return_host else stream
    mark("to_host")
    finish_marks()
    return out, info


class TorchComm:
    """torch.distributed plumbing for scalars and finished pieces (gloo on CPU, nccl on GPUs)."""

    def __init__(self, dist, device=None):
        import torch
        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def _t(self, vals):
        return self.torch.tensor(vals, dtype=self.torch.int64, device=self.device if self.device is not None else "cpu")

    def host_view(self, region, shard_len):
        if hasattr(region, "cpu"):
            # only the edges matter for run_info; fetch lazily from the ends
            return _EdgeView(region, shard_len)
        return region[:shard_len]

    def all_gather_ints(self, vals):
        t = self._t(vals)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[int(x) for x in o.tolist()] for o in out]

    def send_int(self, dst, v):
        self.dist.send(self._t([v]), dst)

    def recv_int(self, src):
        t = self._t([0])
        self.dist.recv(t, src)
        return int(t.item())

    def send_piece(self, dst, piece, nbytes):
        n = (nbytes + 7) & ~3
        if hasattr(piece, "data_ptr"):
            self.dist.send(piece[:n].contiguous(), dst)
        else:
            buf = np.zeros(n, np.uint8)
            buf[: min(n, len(piece))] = np.asarray(piece)[: min(n, len(piece))]
            self.dist.send(self.torch.from_numpy(buf), dst)

    def recv_piece(self, src, nbytes):
        n = (nbytes + 7) & ~3
        t = self.torch.empty(n, dtype=self.torch.uint8, device=self.device if self.device is not None else "cpu")
        self.dist.recv(t, src)
        return t if self.device is not None else t.numpy()


class _EdgeView:
    """numpy-like view of a device tensor that only materialises what run_info touches."""

    def __init__(self, t, n):
        self.t, self.size = t, n

    def __getitem__(self, k):
        if isinstance(k, slice):
            step = k.step or 1
            if step == 1:
                start, stop, _ = k.indices(self.size)
                return self.t[start:stop].cpu().numpy()
            if step == -1 and k.start is None and k.stop is None:
                return _RevView(self.t, self.size)
            raise IndexError(k)
        if k < 0:
            k += self.size
        return int(self.t[k].item())


class _RevView:
    def __init__(self, t, n):
        self.t, self.size = t, n

    def __getitem__(self, k):
        start, stop, _ = k.indices(self.size)
        return self.t[self.size - stop:self.size - start].flip(0).cpu().numpy()


class ThreadComm:
    """In-process stand-in for TorchComm: ranks are threads (tests; one GPU emulating several ranks)."""

    class Shared:
        def __init__(self, world):
            import queue
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world
            self.q = {(a, b): queue.Queue() for a in range(world) for b in range(world)}

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def host_view(self, region, shard_len):
        if hasattr(region, "cpu"):
            return _EdgeView(region, shard_len)
        return region[:shard_len]

    def all_gather_ints(self, vals):
        self.sh.slots[self.rank] = list(vals)
        self.sh.barrier.wait()
        out = [list(v) for v in self.sh.slots]
        self.sh.barrier.wait()
        return out

    def send_int(self, dst, v):
        self.sh.q[(self.rank, dst)].put(int(v))

    def recv_int(self, src):
        return self.sh.q[(src, self.rank)].get(timeout=600)

    def send_piece(self, dst, piece, nbytes):
        n = (nbytes + 7) & ~3
        self.sh.q[(self.rank, dst)].put(piece[:n].clone() if hasattr(piece, "clone") else np.array(piece[:n]))

    def recv_piece(self, src, nbytes):
        return self.sh.q[(src, self.rank)].get(timeout=600)


def run_threads(world, make_backend, shards, halos, level, tail_streamed=False):
    """Runs compress_sharded for `world` ranks as threads; returns rank 0's stream and all infos."""
    import threading
    shared = ThreadComm.Shared(world)
    results = [None] * world
    errors = []

    def work(r):
        try:
            be = make_backend(r)
            region = np.concatenate([shards[r], halos[r]]) if halos[r].size else shards[r]
            ends = sum(s.size for s in shards[r + 1:]) == halos[r].size
            results[r] = compress_sharded(be, ThreadComm(shared, r), be.load(region), int(shards[r].size), level, ends, tail_streamed)
        except Exception as ex:  # noqa: BLE001
            errors.append((r, ex))
            shared.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errors:
        raise errors[0][1]
    return results[0][0], [r[1] for r in results]


def make_halos(shards, halo_bytes):
    """Halo of shard r = the first halo_bytes of everything that follows it."""
    halos = []
    for r in range(len(shards)):
        rest = [s for s in shards[r + 1:]]
        tail = np.concatenate(rest) if rest else np.zeros(0, np.uint8)
        halos.append(np.ascontiguousarray(tail[:halo_bytes]))
    return halos
