"""Where does the host-buffer (kmg_count_ascii) path spend its wall time?  Times count / finalize / histogram
separately for a few chunk sizes on the C4 workload.  Run under gpurun; prints one line per configuration."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import krust_b200 as kb  # noqa: E402
from bench import SEED, expected_windows, workload_slice  # noqa: E402

k, total, records = 21, 3_100_000_000, 31
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
a, b, offsets_np = workload_slice(total, records, 1, 0, k)
exp = expected_windows(total, records, k)
gen = kb.GpuKmerCounter(k, device=0)
d_seq = torch.empty(total + 64, dtype=torch.uint8, device=dev)[:total]
gen.synth_uniform_device(SEED, 0, total, d_seq.data_ptr())
h_seq = torch.empty(total, dtype=torch.uint8, pin_memory=True)
h_seq.copy_(d_seq)
torch.cuda.synchronize()
del d_seq
gen.close()
h_np = h_seq.numpy()

# raw H2D bandwidth for reference
d_tmp = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
d_tmp.copy_(h_seq[: 1 << 30], non_blocking=True)
torch.cuda.synchronize()
print(f"H2D 1 GiB pinned: {(1 << 30) / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
del d_tmp

import os
STREAM = int(os.environ.get("PROBE_STREAM", "0")) or None   # 1 = the legacy default stream (what torch hands out)
for bb in [int(x) for x in (sys.argv[1:] or ["1073741824", "536870912", "268435456", "134217728"])]:
    eng = kb.GpuKmerCounter(k, device=0, expected_distinct=int(exp * 1.03), batch_bases=bb, stream=STREAM)
    best = None
    for it in range(4):
        eng.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.count_batch(h_np, None, offsets_np)
        t1 = time.perf_counter()
        s = eng.finalize(True)
        t2 = time.perf_counter()
        vals, freqs = eng.histogram(1)
        t3 = time.perf_counter()
        row = (t3 - t0, t1 - t0, t2 - t1, t3 - t2)
        if it and (best is None or row[0] < best[0]):
            best = row
            bs = s
    assert int((vals * freqs).sum()) == exp
    print(f"batch_bases={bb:>11}: total {best[0]*1e3:7.1f} ms  count {best[1]*1e3:7.1f}  finalize {best[2]*1e3:7.1f}  hist {best[3]*1e3:6.1f}"
          f"  | scan_ms {bs.get('scan_ns', 0)/1e6:.1f} cons_ms {bs.get('consolidate_ns', 0)/1e6:.1f} kernel_ms {bs.get('kernel_ns', 0)/1e6:.1f}"
          f" grows {bs.get('n_grows')} -> {exp / best[0] / 1e9:.2f} G/s", flush=True)
    eng.close()

if os.environ.get("PROBE_BENCHLIKE"):
    # what bench.py does: device-resident steps first, then host-buffer steps on the SAME engine (pool holds the big blocks)
    from krust_b200.dist import GpuShardEngine
    d_seq = torch.empty(total + 64, dtype=torch.uint8, device=dev)[:total]
    d_seq.copy_(h_seq)
    d_off = torch.from_numpy(offsets_np.astype(np.int64)).to(dev)
    engine = GpuShardEngine(k, dev, expected_distinct=int(exp * 1.03) + 1024)
    for it in range(3):
        engine.reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        engine.count_local(d_seq, d_off)
        engine.finalize(False)
        torch.cuda.synchronize()
        print(f"device step {it}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    for it in range(5):
        t0 = time.perf_counter()
        engine.reset()
        t1 = time.perf_counter()
        engine.counter.count_batch(h_np, None, offsets_np)
        t2 = time.perf_counter()
        engine.finalize(False)
        t3 = time.perf_counter()
        engine.histogram(1)
        t4 = time.perf_counter()
        free_b, total_b = torch.cuda.mem_get_info()
        print(f"host step {it}: reset {(t1-t0)*1e3:.1f} count {(t2-t1)*1e3:.1f} finalize {(t3-t2)*1e3:.1f} hist {(t4-t3)*1e3:.1f} ms; free {free_b/2**30:.1f} GiB", flush=True)
