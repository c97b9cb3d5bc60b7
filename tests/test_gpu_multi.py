"""Multi-GPU parity: the real CUDA engine per rank (one process per GPU), against the oracle -- the fused scatter + exchange
over NVLink (kmg_shard_*, peers mapped through CUDA IPC) and the generic bucket + NCCL all-to-all path.
Needs >= 2 GPUs (skipped on a single-GPU box; the fused path then runs with several ranks on one device in
test_gpu_shard_group.py, the generic plumbing under gloo in test_dist_gloo.py)."""
import os
import pickle
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, k, data_path, out_dir, fused, host_feed):
    import torch
    import torch.distributed as dist
    from krust_b200.dist import GpuShardEngine, ShardedKmerCounter, slice_for_rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["KMG_DIST_FUSED"] = "1" if fused else "0"
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        with open(data_path, "rb") as f:
            seq_np, offsets_np = pickle.load(f)
        total = len(seq_np)
        a, b = slice_for_rank(total, world, rank, k)
        inside = [int(o) - a for o in offsets_np[1:-1] if a < int(o) < b]
        off = torch.tensor([0] + inside + [b - a], dtype=torch.int64, device=dev)
        seq = torch.from_numpy(seq_np[a:b].copy()).to(dev)
        eng = GpuShardEngine(k, dev, batch_bases=400_000)   # several rounds per rank
        sc = ShardedKmerCounter(eng)
        assert sc.fused == bool(fused)
        if host_feed:
            sc.count_host(seq_np[a:b].copy(), off.cpu().numpy().astype(np.uint64))
        else:
            sc.count(seq, off)
        summary = sc.finalize()
        kmix = sc.save_kmix(os.path.join(out_dir, "multi.kmix"))
        keys, counts = sc.export_gathered(1)
        hv, hf = sc.histogram(1)
        with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
            pickle.dump(dict(summary={x: summary[x] for x in ("n_windows", "n_distinct", "max_count")}, keys=keys, counts=counts, hv=hv, hf=hf, kmix=kmix, stats=sc.stats()), f)
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("k,fused,host_feed", [(21, 1, 0), (12, 1, 1), (21, 0, 0), (31, 1, 0)])
def test_multi_gpu_sharded_count_equals_oracle(tmp_path, k, fused, host_feed):
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    world = min(_n_gpus(), 4)
    rng = np.random.default_rng(k)
    seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), p=[.2475, .2475, .2475, .2475, .01], size=2_000_000).astype(np.uint8)
    offsets = np.array([0, 300_000, 300_000, 1_200_000, 2_000_000], dtype=np.uint64)
    data = tmp_path / "d.pkl"
    with open(data, "wb") as f:
        pickle.dump((seq, offsets), f)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, k, str(data), str(tmp_path), fused, host_feed), nprocs=world, join=True)
    okeys, ocounts, windows = orc.count_batch(k, seq, None, offsets, mode="rolling")
    ov, of = orc.histogram(ocounts, 1)
    for r in range(world):
        with open(tmp_path / f"r{r}.pkl", "rb") as f:
            res = pickle.load(f)
        assert (res["keys"] == okeys).all() and (res["counts"] == ocounts).all()
        assert res["summary"]["n_windows"] == windows and res["summary"]["n_distinct"] == len(okeys)
        assert (res["hv"] == ov).all() and (res["hf"] == of).all()
        assert res["kmix"]["records"] == len(okeys)
        if fused:
            assert res["stats"]["rounds"] >= 2 and res["stats"]["sent_keys"] > 0
    kk, ikeys, icounts = orc.kmix_decode(open(tmp_path / "multi.kmix", "rb").read())   # ONE valid index from all GPU shards
    o2 = np.argsort(ikeys, kind="stable")
    assert kk == k and (ikeys[o2] == okeys).all() and (icounts[o2] == ocounts).all()
