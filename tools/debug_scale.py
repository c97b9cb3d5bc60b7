import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krust_b200 as kb
from krust_b200 import _lib
PART=_lib.KMG_FLAG_FORCE_PARTITIONED
dev=torch.device("cuda:0"); torch.cuda.set_device(dev)
n=200_000_000; k=21
buf=torch.empty(n,dtype=torch.uint8,device=dev)
stream=torch.cuda.current_stream().cuda_stream
with kb.GpuKmerCounter(k, stream=stream, expected_distinct=n) as c:
    c.synth_uniform_device(44, 0, n, buf.data_ptr())
    c.count_device(buf.data_ptr(), n); s=c.finalize(); print("direct scan:", s["n_windows"], s["n_distinct"], s["path"], flush=True)
    ref_keys, ref_counts = c.export(1, True)
for shards in (1, 2):
    out=torch.empty(n,dtype=torch.int64,device=dev)
    with kb.GpuKmerCounter(k, stream=stream) as c:
        counts=c.extract_keys_device(buf.data_ptr(), n, shards, out.data_ptr(), n)
    tot=int(counts.sum()); print("extract shards",shards,counts, flush=True)
    host=out[:tot].cpu().numpy().view(np.uint64)
    print("   distinct in extracted keys:", len(np.unique(host)), flush=True)
    with kb.GpuKmerCounter(k, stream=stream, expected_distinct=n, flags=PART) as c:
        c.insert_keys_device(out.data_ptr(), tot); s=c.finalize(); print("   insert all:", s["n_windows"], s["n_distinct"], flush=True)
        gk,gc=c.export(1,True)
        print("   equal to direct:", len(gk)==len(ref_keys) and bool((gk==ref_keys).all() and (gc==ref_counts).all()), flush=True)
