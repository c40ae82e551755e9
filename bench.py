#!/usr/bin/env python
"""bench.py -- compress MB/s at -9 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload text|random|period1000|aab|runs|mixed]
                  [--mb 1000] [--level 9]

A "step" is one pass of the whole compression path (RLE1+CRC -> BWT -> MTF/RLE2 -> Huffman -> stream)
over one synthetic input of --mb MB (default: the 1 GB Zipf text of SURVEY.md 8(d) C2, BASELINE.json configs[1]).

  value  MB/s of input, input and output resident in HBM (bz2b200_compress_device), CUDA events on the
         stream the kernels run on, max over ranks
  e2e    the same metric through BZ2_bzBuffToBuffCompress with pinned HOST buffers (H2D + D2H inside)
  roofline      dominant stage (BWT) and whole pass against the measured HBM copy bandwidth
  cpu_baseline  the reference's own CPU path (oracle/_ref) on a bounded sample, single thread

With --impl reference the reference CPU implementation (oracle/_ref, else the oracle port) is timed on all
host threads on a bounded sample of the same workload.  N>1: one process per GPU (torchrun), each rank
compresses its own shard of the same size (weak scaling, no collective on the data path).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def make_input(workload, n, rank=0):
    import support as S
    seed = S.TEXT_SEED + rank
    if workload == "text":
        return S.gen_text(n, seed=seed)
    if workload == "random":
        return S.gen_random(n, seed=2 + rank)
    if workload == "period1000":
        return S.gen_period1000(n)
    if workload == "aab":
        return S.gen_tile(n, b"aab")
    if workload == "runs":
        return S.gen_runs(n, seed=3 + rank)
    if workload == "mixed":
        return S.gen_c4(n, seg=64 << 20)
    raise SystemExit(f"unknown workload {workload}")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line) through NVML,
    at a rate that adapts to what a query costs: a query takes the driver's lock, and on some boxes spawning
    nvidia-smi (or polling NVML every few ms) stretched the timed step by 10 % to 4x."""

    def __init__(self, index, count=1, mode=None):
        self.rows = []
        self.stop = False
        self.index = index
        self.count = count            # GPUs index .. index+count-1 are sampled
        self.how = "nvml"
        self.query_ms = []
        # proc (default): a small helper process polls NVML (0.02 ms per query, no effect on the step).  thread: a sampler
        # thread in this process -- its queries sometimes waited 20-290 ms behind the process's own CUDA calls and stalled
        # one timed step in three at N = 2 (9.1 instead of 11.3 GB/s).  inline: the timed loop polls between steps.
        self.mode = mode or os.environ.get("BENCH_CLOCKS", "proc")
        self.th = threading.Thread(target=self.run, daemon=True)
        self.proc = None
        self._nv = None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        hs = []
        for k in range(self.index, self.index + self.count):
            idx = k
            if vis:
                try:
                    idx = int(vis.split(",")[k])
                except Exception:  # noqa: BLE001
                    pass
            hs.append(nv.nvmlDeviceGetHandleByIndex(idx))
        mx = [nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM) for h in hs]
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, 0), (nv.nvmlClocksEventReasonHwThermalSlowdown, 1),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, 2), (nv.nvmlClocksEventReasonSwPowerCap, 3)]
        self.ready.set()
        while not self.stop:
            t0 = time.perf_counter()
            for h, m in zip(hs, mx):
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                row = [str(sm), str(m)] + ["Not Active"] * 4
                for bit, k in bits:
                    if r & bit:
                        row[2 + k] = "Active"
                self.rows.append(row)
            dt = time.perf_counter() - t0
            self.query_ms.append(dt * 1e3)
            time.sleep(max(0.1, 30.0 * dt))     # keep the sampler under ~3 % of the wall clock

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(1.0)

    def run(self):
        try:
            self._run_nvml()
        except Exception:  # noqa: BLE001
            self.how = "nvidia-smi"
            self.ready.set()
            self._run_smi()

    def _handles(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        hs = []
        for k in range(self.index, self.index + self.count):
            idx = k
            if vis:
                try:
                    idx = int(vis.split(",")[k])
                except Exception:  # noqa: BLE001
                    pass
            hs.append(nv.nvmlDeviceGetHandleByIndex(idx))
        return nv, hs

    def poll(self):
        """inline mode: one sample, taken by the caller's thread between two steps of the timed region"""
        if self.mode != "inline" or self._nv is None:
            return
        nv, hs, mx = self._nv
        t0 = time.perf_counter()
        for h, m in zip(hs, mx):
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            row = [str(sm), str(m)] + ["Not Active"] * 4
            for bit, k in ((nv.nvmlClocksEventReasonHwSlowdown, 0), (nv.nvmlClocksEventReasonHwThermalSlowdown, 1),
                           (nv.nvmlClocksEventReasonSwThermalSlowdown, 2), (nv.nvmlClocksEventReasonSwPowerCap, 3)):
                if r & bit:
                    row[2 + k] = "Active"
            self.rows.append(row)
        self.query_ms.append((time.perf_counter() - t0) * 1e3)

    def __enter__(self):
        if self.mode == "inline":
            try:
                nv, hs = self._handles()
                self._nv = (nv, hs, [nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM) for h in hs])
                self.how = "nvml, sampled by the timing thread between steps"
                return self
            except Exception:  # noqa: BLE001
                self.mode = "thread"
        if self.mode == "proc":
            code = ("import sys,time,pynvml as nv\nnv.nvmlInit()\nh=nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))\n"
                    "m=nv.nvmlDeviceGetMaxClockInfo(h,nv.NVML_CLOCK_SM)\nprint('ready',flush=True)\n"
                    "while True:\n t=time.perf_counter()\n s=nv.nvmlDeviceGetClockInfo(h,nv.NVML_CLOCK_SM)\n"
                    " r=nv.nvmlDeviceGetCurrentClocksEventReasons(h)\n d=(time.perf_counter()-t)*1e3\n"
                    " print(s,m,r,d,flush=True)\n time.sleep(max(0.1,30*d/1e3))\n")
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:  # noqa: BLE001
                    pass
            try:
                self.proc = subprocess.Popen([sys.executable, "-c", code, str(idx)], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)
                if self.proc.stdout.readline().strip() != "ready":      # NVML did not come up in the helper
                    raise RuntimeError("clock helper failed")
                self.how = "nvml, helper process"
                return self
            except Exception:  # noqa: BLE001
                if self.proc is not None:
                    self.proc.kill()
                self.proc = None
                self.mode = "thread"
        self.ready = threading.Event()
        self.th.start()
        self.ready.wait(timeout=10)          # NVML initialisation stays outside the timed region
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:  # noqa: BLE001
                out = ""
            for ln in out.splitlines():
                f = ln.split()
                if len(f) == 4:
                    r = int(f[2])
                    self.rows.append([f[0], f[1]] + ["Active" if r & b else "Not Active" for b in (0x8, 0x40, 0x20, 0x4)])
                    self.query_ms.append(float(f[3]))
            return
        if self.mode == "inline":
            return
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max(int(r[1]) for r in self.rows if r[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows if len(r) > 2 + k)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows),
                "source": self.how, "query_ms": round(sorted(self.query_ms)[len(self.query_ms) // 2], 3) if self.query_ms else None}


def cpu_reference_rate(data, level, threads, seconds_budget, sample_bytes):
    """Reference CPU path on `threads` host threads, each compressing its own contiguous slice of a bounded
    sample (pbzip2-style independent streams).  Returns (MB/s, kind, sample description)."""
    import support as S
    from concurrent.futures import ThreadPoolExecutor
    use_ref = S.have_ref()
    lib = S.ref() if use_ref else S.oracle()
    sample = data[: min(sample_bytes, data.size)]
    per = sample.size // threads
    outs = [np.empty(int(per * 1.02) + 70000, np.uint8) for _ in range(threads)]

    def work(t):
        sl = sample[t * per:(t + 1) * per]
        if use_ref:
            n = C.c_uint(outs[t].size)
            rc = lib.BZ2_bzBuffToBuffCompress(S._p(outs[t]), C.byref(n), S._p(sl), sl.size, level, 0, 0)
            assert rc == 0
            return n.value
        return lib.orc_compress(S._p(sl), sl.size, level, 1, None, S._p(outs[t]), outs[t].size)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    kind = "reference" if use_ref else "port"
    return per * threads / dt / 1e6, kind, f"{per * threads} B prefix of the workload as {threads} independent slice(s), -{level}, one pass, {dt:.1f} s"


def merge_clocks(per_rank):
    """One clocks object for the node: the lowest per-GPU median SM clock, every throttle reason any GPU reported."""
    ok = [c for c in per_rank if c and c.get("sm_mhz") is not None]
    if not ok:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    reasons = sorted({r for c in ok for r in c.get("reasons", [])})
    return {"sm_mhz": min(c["sm_mhz"] for c in ok), "sm_max_mhz": max(c["sm_max_mhz"] for c in ok), "reasons": reasons,
            "samples": sum(c.get("samples", 0) for c in ok), "source": ok[0].get("source"),
            "query_ms": max((c.get("query_ms") or 0.0) for c in ok), "per_gpu_sm_mhz": [c["sm_mhz"] for c in ok]}


def sharded_bench(args, torch, dist, B, rank, world, local_rank, dev, data, n, metric, config):
    """N > 1: ONE .bz2 stream, sharded by block across the ranks (bzip2_b200/sharding.py): every rank scans
    its shard (+ halo) for chunk ends, the block-boundary chain is one integer handed rank to rank, every rank
    compresses its block-aligned segment, rank 0 bit-shifts the pieces into place over NVLink (S5)."""
    from bzip2_b200 import sharding as sh
    halo_bytes = min(n, 64 << 20)
    halo = make_input(args.workload, halo_bytes, rank + 1) if rank + 1 < world else np.zeros(0, np.uint8)
    region_h = torch.from_numpy(np.concatenate([data, halo])).pin_memory()
    be = sh.GpuBackend(args.level, local_rank)
    comm = sh.TorchComm(dist, dev)
    ends = rank == world - 1
    region_d = region_h.to(dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    h2d_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def one(resident):
        if resident:
            reg = region_d
        else:
            h2d_ev[0].record()
            reg = region_h.to(dev, non_blocking=True)
            h2d_ev[1].record()
        out, info = sh.compress_sharded(be, comm, reg, n, args.level, ends, return_host=not resident)
        if not resident:
            info["protocol_ms"]["h2d_copy"] = round(h2d_ev[0].elapsed_time(h2d_ev[1]), 3)
        return out, info

    # the engine, the NCCL plumbing and the timing events all sit on torch's current stream of this device
    be.eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    per_step = []          # this rank's host-side time of every timed step (diagnostic)

    def timed(resident, steps, poll=None):
        """K steps bracketed by barrier + synchronize, timed on the device; returns (max over ranks in s, last result)."""
        barrier()
        ev0.record()
        t0 = time.perf_counter()
        per_step.clear()
        for _ in range(steps):
            t1 = time.perf_counter()
            res = one(resident)
            per_step.append(round((time.perf_counter() - t1) * 1e3, 2))
            if poll is not None:
                poll()
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([ev0.elapsed_time(ev1) * 1e-3, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0].item()), float(tt[1].item()), res

    for _ in range(args.warmup):
        out, info = one(True)
    # every rank samples its own GPU (helper process, see ClockSampler); rank 0 merges the summaries
    with ClockSampler(local_rank, 1) as clk:
        dev_s, wall_s, (out, info) = timed(True, args.steps, clk.poll)
    clk_all = [None] * world
    dist.all_gather_object(clk_all, clk.summary())
    ms_per_step = dev_s / args.steps * 1e3
    resident_steps = list(per_step)
    proto = [None] * world
    dist.all_gather_object(proto, info.get("protocol_ms"))
    value = world * n / (ms_per_step * 1e-3) / 1e6
    e2e = None
    if not args.no_e2e:
        for _ in range(max(1, args.warmup)):
            host_out, info = one(False)
        e_dev_s, e_wall_s, (host_out, info) = timed(False, args.steps)
        e_proto = [None] * world
        dist.all_gather_object(e_proto, info.get("protocol_ms"))
        e2e = {"value": round(world * n * args.steps / max(e_dev_s, e_wall_s) / 1e6, 2), "unit": "MB/s", "protocol_ms_per_rank": e_proto, "step_ms_rank0": list(per_step),
               "h2d_bytes_per_step": int(region_h.numel()), "d2h_bytes_per_step": int(info["total_bytes"]) if rank == 0 else 0,
               "api": "bzip2_b200.sharding.compress_sharded over bz2b200_scan_* / bz2b200_compress_device / bz2b200_concat_bits, pinned host buffers"}
        if rank == 0:
            import bz2 as _bz2
            # the assembled stream is one valid .bz2 stream: decode its head with an independent decoder
            d = _bz2.BZ2Decompressor()
            head = d.decompress(bytes(host_out[: 4 << 20]), max_length=1 << 20)
            assert head == data[: len(head)].tobytes(), "sharded stream does not decode to the input"
    st = be.eng.stats
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        seg_in = info["segment"][1] - info["segment"][0]
        s2_bytes = 2.0 * st.sum_nblock
        roof = {"bound": "hbm", "kernel": "S2 BWT stage of rank 0 (k-gram bucket + prefix-doubling kernels)",
                "achieved": round(s2_bytes / (st.ms_s2 * 1e-3) / 1e9, 3), "peak": peak, "unit": "GB/s",
                "frac": round(s2_bytes / (st.ms_s2 * 1e-3) / 1e9 / peak, 6), "traffic": None,
                "stage_ms_rank0": {"s1": round(st.ms_s1, 3), "s2": round(st.ms_s2, 3), "s3": round(st.ms_s3, 3), "s4": round(st.ms_s4, 3)},
                "rank0_segment_bytes": int(seg_in)}
        config = dict(config)
        config["workload"] += f"; ONE stream of {world} x {args.mb} MB sharded by block, 64 MiB halo per rank"
        line = {"metric": metric, "value": round(value, 2), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config, "clocks": merge_clocks(clk_all),
                "e2e": e2e, "gpu_launches": int(st.kernel_launches), "roofline": roof, "cpu_baseline": None,
                "out_bytes": int(info["total_bytes"]), "blocks_rank0": int(st.n_blocks),
                "wall_ms_per_step": round(wall_s / args.steps * 1e3, 3), "protocol_ms_per_rank": proto, "step_ms_rank0": resident_steps}
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="text")
    ap.add_argument("--mb", type=int, default=1000)
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.mb * 1_000_000
    metric = "compress MB/s at -9 (1/2/4/8 B200) vs host libbz2; byte-exact .bz2 output"
    config = {"workload": f"{args.mb} MB synthetic {args.workload} (SURVEY 8d; xorshift64* seeded), blockSize100k={args.level}, "
                          f"one shard per GPU", "l2": "input (>= 1 GB) exceeds the 126 MB L2; no flush needed",
              "level": args.level, "bytes_per_gpu": n}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        sample_bytes = min(n, threads * 24_000_000)
        data = make_input(args.workload, sample_bytes)
        vals = []
        info = None
        for it in range(args.warmup + args.steps):
            v, kind, sample = cpu_reference_rate(data, args.level, threads, 30, sample_bytes)
            if it >= args.warmup:
                vals.append(v)
            info = (kind, sample)
        v = sum(vals) / len(vals)
        line = {"impl": "reference", "metric": metric, "value": round(v, 2), "unit": "MB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sample_bytes / (v * 1e6) * 1e3, 2),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": round(v, 2), "unit": "MB/s", "cores": threads, "kind": info[0], "sample": info[1]},
                "e2e": {"value": round(v, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import bzip2_b200 as B
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    os.environ["BZ2_B200_DEVICE"] = str(local_rank)      # device used by the libbz2 entry points
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    data = make_input(args.workload, n, rank)
    if world > 1:
        return sharded_bench(args, torch, dist, B, rank, world, local_rank, dev, data, n, metric, config)
    h_in = torch.from_numpy(data).pin_memory()
    d_in = h_in.to(dev)
    cap = n + n // 50 + 24576 * (n // (100000 * args.level - 19) + 2) + 1024
    cap = (cap + 255) & ~255
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()

    eng = B.Engine(level=args.level, device=local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    out_len = 0
    for _ in range(args.warmup):
        out_len = eng.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), cap)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(5)
    launches = 0
    with ClockSampler(local_rank) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            out_len = eng.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), cap)
            st = eng.stats
            stage_ms += np.array([st.ms_total, st.ms_s1, st.ms_s2, st.ms_s3, st.ms_s4])
            launches += st.kernel_launches
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    st = eng.stats
    stage_ms /= args.steps

    # end to end through the libbz2 entry point with host buffers
    e2e = None
    if not args.no_e2e:
        eng.set_stream(None)
        lib = B.load()
        # BZ2_bzBuffToBuffCompress takes 32-bit lengths; the engine pool keeps the HBM allocation between calls
        dlen = C.c_uint(min(cap, 0xFFFFFFFF))
        for _ in range(max(1, args.warmup)):
            dlen = C.c_uint(min(cap, 0xFFFFFFFF))
            rc = lib.BZ2_bzBuffToBuffCompress(h_out.data_ptr(), C.byref(dlen), h_in.data_ptr(), n, args.level, 0, 0)
            assert rc == 0, rc
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dlen = C.c_uint(min(cap, 0xFFFFFFFF))
            rc = lib.BZ2_bzBuffToBuffCompress(h_out.data_ptr(), C.byref(dlen), h_in.data_ptr(), n, args.level, 0, 0)
            assert rc == 0, rc
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": round(world * n * args.steps / dt / 1e6, 2), "unit": "MB/s", "h2d_bytes_per_step": n,
               "d2h_bytes_per_step": int(dlen.value), "api": "BZ2_bzBuffToBuffCompress, pinned host buffers"}
        # the host path must give the same bytes as the device path
        same = bytes(h_out[: dlen.value].numpy()[:4096]) == bytes(d_out[:4096].cpu().numpy())
        assert dlen.value == out_len and same, "host and device paths disagree"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    rho = st.sum_nblock / max(1, st.in_bytes)
    mu = st.sum_nmtf / max(1, st.sum_nblock)
    c = st.out_bytes / max(1, st.in_bytes)
    A = 1 + 4 * rho + 12 * rho * mu + 3 * c              # algorithmic bytes per input byte, whole pass
    s2_bytes = 2 * rho * n                               # BWT stage: read block once, write last column once
    # DRAM traffic of the same stage from the committed ncu capture (profiles/r01_traffic.json), scaled to this step
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if args.workload == "text" and args.level == 9:
            traffic = int(tr["per_input_byte"]["S2"]["dram_bytes"] * n)
    except Exception:  # noqa: BLE001
        pass
    roof = {"bound": "hbm",
            "kernel": "S2 BWT kernel family (k_kgram*, k_refine_*, k_resolve_periodic, k_rep_*, k_bwt_out; ~70 launches per 100 MB window), timed live by CUDA events "
                      "around the stage on the launching stream; top single kernel k_refine_large<true> = 10% of the step "
                      "(profiles/r01_v6_launch_summary.md)",
            "algorithmic_bytes": int(s2_bytes), "algorithmic_rule": "2*rho bytes per input byte: read the block once, write the last column once (SURVEY 8d)",
            "achieved": round(s2_bytes / (stage_ms[2] * 1e-3) / 1e9, 3), "peak": peak, "unit": "GB/s",
            "frac": round(s2_bytes / (stage_ms[2] * 1e-3) / 1e9 / peak, 6), "traffic": traffic, "peak_source": peak_src,
            "whole_pass": {"A_bytes_per_input_byte": round(A, 3), "rho": round(rho, 4), "mu": round(mu, 4), "c": round(c, 4),
                           "achieved": round(A * n / (stage_ms[0] * 1e-3) / 1e9, 3),
                           "frac": round(A * n / (stage_ms[0] * 1e-3) / 1e9 / peak, 6)},
            "stage_ms": {"s1_rle_crc": round(float(stage_ms[1]), 3), "s2_bwt": round(float(stage_ms[2]), 3),
                         "s3_mtf": round(float(stage_ms[3]), 3), "s4_huffman_pack": round(float(stage_ms[4]), 3),
                         "sum_events": round(float(stage_ms[0]), 3)}}
    cpu = None
    if not args.no_cpu:
        v, kind, sample = cpu_reference_rate(data, args.level, 1, 20, 200_000_000)
        cpu = {"value": round(v, 2), "unit": "MB/s", "cores": 1, "kind": kind, "sample": sample}
    line = {"metric": metric, "value": round(value, 2), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config, "clocks": clk.summary(),
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "out_bytes": int(out_len), "blocks": int(st.n_blocks), "bwt_rounds": int(st.bwt_rounds)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
