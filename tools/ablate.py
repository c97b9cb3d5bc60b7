"""Ablation timing of the legacy warp-scatter A1 kernel (select it with KMG_SCATTER=1): KMG_DEBUG=bits python tools/ablate.py
(results are WRONG with bits set).  The default rows kernel has no ablation switches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krust_b200 as kb
dev=torch.device("cuda:0"); n=1_000_000_000; k=21
buf=torch.empty(n,dtype=torch.uint8,device=dev)
out=torch.empty(n,dtype=torch.int64,device=dev)
with kb.GpuKmerCounter(k) as c:
    c.synth_uniform_device(44,0,n,buf.data_ptr())
    for bins in (928, 64):
        for rep in range(2):
            torch.cuda.synchronize(); t0=time.perf_counter()
            try:
                c.extract_keys_device(buf.data_ptr(), n, bins, out.data_ptr(), n)
            except Exception as e:
                print("err", str(e)[:80])
            torch.cuda.synchronize(); dt=time.perf_counter()-t0
        print(f"KMG_DEBUG={os.environ.get('KMG_DEBUG','0')} bins={bins}: extract(count+scatter) {dt*1e3:.1f} ms for {n/1e9:.1f}e9 bases", flush=True)
