import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krust_b200 as kb
from krust_b200 import _lib
from oracle import oracle as orc
recs=[b"ACGTACGT", b"GATTACA"]; k=4
okeys,ocounts,_=orc.count_records(k,recs)
print("oracle", [(orc.unpack(int(a),k).decode(),int(b)) for a,b in zip(okeys,ocounts)])
dev=torch.device("cuda:0")
seq=torch.from_numpy(np.frombuffer(b"".join(recs),dtype=np.uint8).copy()).to(dev)
off=torch.tensor([0,8,15],dtype=torch.int64,device=dev)
for P in (16,32,64):
    out=torch.zeros(64,dtype=torch.int64,device=dev)
    with kb.GpuKmerCounter(k) as c:
        cnt=c.extract_keys_device(seq.data_ptr(),15,P,out.data_ptr(),64,d_offsets=off.data_ptr(),n_records=2)
    keys=out[:int(cnt.sum())].cpu().numpy().view(np.uint64)
    print("extract P",P,int(cnt.sum()),sorted(orc.unpack(int(a),k).decode() for a in keys))
for pl in (3,4,5,6):
    with kb.GpuKmerCounter(k,flags=_lib.KMG_FLAG_FORCE_PARTITIONED,parts_log2=pl) as c:
        c.count_records(recs); s=c.finalize(); keys,counts=c.export(1,True)
    print("part pl",pl,s["n_windows"],s["n_distinct"],[(orc.unpack(int(a),k).decode(),int(b)) for a,b in zip(keys,counts)])
    with kb.GpuKmerCounter(k,flags=_lib.KMG_FLAG_FORCE_PARTITIONED,parts_log2=pl) as c:
        c.count_records(recs); s=c.finalize(); keys,counts=c.export(1,False)
    print("   unsorted",[(orc.unpack(int(a),k).decode(),int(b)) for a,b in zip(keys,counts)])
