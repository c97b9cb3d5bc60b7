import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krust_b200 as kb
from krust_b200 import _lib
from oracle import oracle as orc
PART=_lib.KMG_FLAG_FORCE_PARTITIONED
dev=torch.device("cuda:0")
rng=np.random.default_rng(1)
def cmp(tag, got, exp):
    gk,gc=got; ek,ec=exp
    ok = len(gk)==len(ek) and (gk==ek).all() and (gc==ec).all()
    print(tag, "OK" if ok else f"FAIL got {len(gk)} distinct sum {int(gc.sum())} expected {len(ek)} sum {int(ec.sum())}", flush=True)
    if not ok and len(gk):
        print("   first got", gk[:5], gc[:5], " first exp", ek[:5], ec[:5])
k=11
keys = rng.integers(0, 4**k, size=300_000, dtype=np.uint64)
uk,uc = np.unique(keys, return_counts=True); uc=uc.astype(np.uint64)
t = torch.from_numpy(keys.view(np.int64)).to(dev)
for pl in (5, 3, 8):
    with kb.GpuKmerCounter(k, flags=PART, parts_log2=pl) as c:
        c.insert_keys_device(t.data_ptr(), len(keys)); c.finalize(); cmp(f"pl{pl} keys R=1", c.export(1,True), (uk,uc))
    with kb.GpuKmerCounter(k, flags=PART, parts_log2=pl) as c:
        h=len(keys)//2
        c.insert_keys_device(t.data_ptr(), h); c.insert_keys_device(t[h:].data_ptr(), len(keys)-h); c.finalize(); cmp(f"pl{pl} keys R=2", c.export(1,True), (uk,uc))
    with kb.GpuKmerCounter(k, flags=PART, parts_log2=pl) as c:
        h=len(keys)//2
        rk,rc=np.unique(keys[h:],return_counts=True)
        tk=torch.from_numpy(rk.view(np.int64)).to(dev); tc=torch.from_numpy(rc.astype(np.int64)).to(dev)
        c.insert_keys_device(t.data_ptr(), h); c.insert_keys_device(tk.data_ptr(), len(rk), tc.data_ptr()); c.finalize(); cmp(f"pl{pl} keys+pairs R=2", c.export(1,True), (uk,uc))
    with kb.GpuKmerCounter(k, flags=PART, parts_log2=pl) as c:
        rk,rc=np.unique(keys,return_counts=True)
        tk=torch.from_numpy(rk.view(np.int64)).to(dev); tc=torch.from_numpy(rc.astype(np.int64)).to(dev)
        c.insert_keys_device(tk.data_ptr(), len(rk), tc.data_ptr()); c.finalize(); cmp(f"pl{pl} pairs R=1", c.export(1,True), (uk,uc))
seq = rng.choice(np.frombuffer(b"ACGT",dtype=np.uint8), size=200_000).astype(np.uint8)
for kk in (11, 21):
    exp = orc.count_batch(kk, seq, None, np.array([0,len(seq)],dtype=np.uint64))
    for bb in (0, 70_000, 20_000, 8192):
        with kb.GpuKmerCounter(kk, flags=PART, parts_log2=5, batch_bases=bb) as c:
            c.count_batch(seq, None, np.array([0,len(seq)],dtype=np.uint64)); s=c.finalize(); cmp(f"k{kk} scan bb={bb} cons={s['n_grows']}", c.export(1,True), exp[:2])
