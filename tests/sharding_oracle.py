"""CPU checker backend for the multi-process sharding logic of bzip2_b200/sharding.py (tests only).

It implements the backend interface of that module with the oracle (oracle/liboracle.so), so that the host logic --
run_info / prev_state / the boundary chain / the bit-level assembly -- can be exercised under gloo without a GPU."""
import ctypes as C

import numpy as np

from bzip2_b200.sharding import TAIL_STREAMED, or_bits


class OracleBackend:
    """CPU checker backend (tests only)."""

    def __init__(self, level):
        import support as S
        self.S, self.level = S, level

    def load(self, region):
        return np.ascontiguousarray(region, np.uint8)

    def scan(self, region, prev_byte, prev_run, input_ends):
        return (region, input_ends)

    def boundary(self, scan, start, limit, tail_streamed):
        region, input_ends = scan
        S = self.S
        lib = S.oracle()
        lib.orc_find_boundary.restype = C.c_uint64
        lib.orc_find_boundary.argtypes = [S.u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
        nb = C.c_uint32(0)
        b = lib.orc_find_boundary(S._buf(region), region.size, start, min(limit, region.size), self.level,
                                  0 if tail_streamed else 1, 1 if input_ends else 0, C.byref(nb))
        if b == 0xFFFFFFFFFFFFFFFF:
            raise RuntimeError("halo too small for the block chain")
        return int(b), int(nb.value)

    def compress_segment(self, region, start, end, flags):
        S = self.S
        lib = S.oracle()
        lib.orc_compress_ex.restype = C.c_int64
        lib.orc_compress_ex.argtypes = [S.u8p, C.c_uint64, C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_int32), S.u8p, C.c_uint64,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        seg = np.ascontiguousarray(region[start:end])
        cap = int(seg.size * 1.3) + 100000
        out = np.zeros(cap, np.uint8)
        bits, fold, nblk = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        n = lib.orc_compress_ex(S._buf(seg), seg.size, self.level, 0 if (flags & TAIL_STREAMED) else 1, flags & 6, None,
                                S._p(out), cap, C.byref(bits), C.byref(fold), C.byref(nblk))
        assert n >= 0, n
        return out[: (bits.value + 7) // 8 + 8], int(bits.value), int(nblk.value), int(fold.value)

    def new_stream(self, nbytes):
        return np.zeros(nbytes + 8, np.uint8)

    def place(self, stream, bit, piece, nbits):
        if isinstance(piece, (bytes, bytearray)):
            piece = np.frombuffer(bytes(piece), np.uint8)
        or_bits(stream, bit, np.asarray(piece), nbits)

    def to_host(self, stream, nbytes):
        return stream[:nbytes].tobytes()
