import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import support as S
from bzip2_b200 import sharding as sh
import bzip2_b200 as B
n=1_000_000_000
data=S.gen_text(n)
eng=B.Engine(level=9)
d_in=torch.from_numpy(data).cuda(); cap=n+n//50+24576*1200
d_out=torch.empty(cap,dtype=torch.uint8,device='cuda')
for _ in range(2): eng.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), cap)
t=time.time(); L=eng.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), cap); torch.cuda.synchronize(); print('single', time.time()-t, L)
single=bytes(d_out[:L].cpu().numpy())
eng.close()
import threading
for world in (2,3,4):
    shards=[np.ascontiguousarray(data[r*n//world:(r+1)*n//world]) for r in range(world)]
    halos=sh.make_halos(shards, 8<<20)
    bes={}
    regions={}
    for r in range(world):
        bes[r]=sh.GpuBackend(9,0)
        regions[r]=bes[r].load(np.concatenate([shards[r],halos[r]]))
    def runonce():
        shared=sh.ThreadComm.Shared(world); res=[None]*world
        def work(r):
            ends = (r==world-1)
            res[r]=sh.compress_sharded(bes[r], sh.ThreadComm(shared,r), regions[r], int(shards[r].size), 9, ends, return_host=False)
        th=[threading.Thread(target=work,args=(r,)) for r in range(world)]
        [x.start() for x in th]; [x.join() for x in th]
        return res
    runonce(); torch.cuda.synchronize()
    t=time.time(); res=runonce(); torch.cuda.synchronize(); dt=time.time()-t
    out=res[0][0]; nb=res[0][1]['total_bytes']
    ok = bytes(out[:nb].cpu().numpy())==single
    print('world',world,'time',round(dt,4),'GB/s',round(n/dt/1e9,2),'ok',ok)
    for b in bes.values(): b.eng.close()
