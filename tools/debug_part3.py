import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krust_b200 as kb
from krust_b200 import _lib
PART=_lib.KMG_FLAG_FORCE_PARTITIONED
dev=torch.device("cuda:0")
rng=np.random.default_rng(1)
k=11
keys = rng.integers(0, 4**k, size=20_000, dtype=np.uint64)
t = torch.from_numpy(keys.view(np.int64)).to(dev)
torch.cuda.synchronize()
with kb.GpuKmerCounter(k, flags=PART, parts_log2=5) as c:
    c.insert_keys_device(t.data_ptr(), len(keys)); c.finalize(); gk,gc=c.export(1,True)
uk,uc=np.unique(keys,return_counts=True)
print(len(gk), len(uk))
