"""Scratch diagnostic: where does host time go inside compress_device when NCCL is up? (torchrun, 2 ranks)"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import support as S
from bzip2_b200 import binding as B, sharding as sh

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
os.environ["BZ2_B200_DEVICE"] = str(lr)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
n = 1000_000_000
data = S.gen_text(n, seed=S.TEXT_SEED + rank)
d = torch.from_numpy(data).to(dev)
be = sh.GpuBackend(9, lr)

def t_compress(tag):
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        be.compress_segment(d, 0, n, 6)
        ts.append((time.perf_counter() - t0) * 1e3)
    st = be.eng.stats
    print(f"[rank {rank}] {tag}: compress_segment ms = {[round(x,1) for x in ts]}  stage sum = {st.ms_s1+st.ms_s2+st.ms_s3+st.ms_s4:.1f}", flush=True)

t_compress("before init_process_group")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    t_compress("after init_process_group")
    x = torch.ones(4, device=dev); dist.all_reduce(x); torch.cuda.synchronize()
    t_compress("after first all_reduce")
    if rank == 0: dist.send(x, 1)
    else: dist.recv(x, 0)
    torch.cuda.synchronize()
    t_compress("after first send/recv")
    h = be.scan(d, 256, 0, True); be.free_scan(h)
    t_compress("after scan create/destroy")
    dist.destroy_process_group()
