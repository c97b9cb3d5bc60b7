"""world_size-2 (and 3) tests of the multi-GPU host logic on CPU with the gloo backend.

The exchange plumbing under test is the product's (krust_b200.dist.ShardedKmerCounter,
slice_for_rank, merge_histograms, kmg_owner_of); only the per-rank engine is a CPU stand-in built on
the oracle, because the CUDA engine cannot run in this container.  The same flow with the real
engine is exercised on GPUs by tests/test_gpu_multi.py.
"""
import os
import pickle
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import krust_b200 as kb
from krust_b200.dist import ShardedKmerCounter, merge_histograms, slice_for_rank
from oracle import oracle as orc


def np_owner_of(keys: np.ndarray, n: int) -> np.ndarray:
    """numpy restatement of kmg::part_of (krust_b200/csrc/kmg_device.cuh) for the stand-in engine."""
    x = keys.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
        h = x >> np.uint64(32)
        return ((h * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


class OracleShardEngine:
    """CPU stand-in with GpuShardEngine's duck-typed interface (test infrastructure only)."""
    device = torch.device("cpu")

    def __init__(self, k, min_quality=None):
        self.k, self.q = k, min_quality
        self.keys = np.zeros(0, dtype=np.uint64)
        self.counts = np.zeros(0, dtype=np.uint64)
        self.n_records = self.n_bases = 0

    def _scan(self, seq, offsets, qual):
        s = seq.numpy()
        off = np.array([0, len(s)], dtype=np.uint64) if offsets is None else offsets.numpy().astype(np.uint64)
        self.n_records += len(off) - 1; self.n_bases += len(s)
        keys, counts, _ = orc.count_batch(self.k, s, None if qual is None else qual.numpy(), off, self.q)
        return keys, counts

    def _add(self, keys, counts):
        allk = np.concatenate([self.keys, keys]); allc = np.concatenate([self.counts, counts])
        self.keys, inv = np.unique(allk, return_inverse=True)
        self.counts = np.bincount(inv, weights=allc.astype(np.float64), minlength=len(self.keys)).astype(np.uint64)

    def count_local(self, seq, offsets=None, qual=None):
        self._add(*self._scan(seq, offsets, qual))

    def extract(self, seq, n_shards, offsets=None, qual=None):
        keys, counts = self._scan(seq, offsets, qual)
        flat = np.repeat(keys, counts.astype(np.int64))
        owner = np_owner_of(flat, n_shards)
        order = np.argsort(owner, kind="stable")
        return torch.from_numpy(flat[order].view(np.int64)), np.bincount(owner, minlength=n_shards).astype(np.uint64)

    def plan(self, expected_keys):
        return 5  # any common coarse-bin count works for the stand-in

    def adopt(self, keys, bin_counts):
        k = keys.numpy().view(np.uint64)
        assert int(np.asarray(bin_counts).sum()) == len(k)
        # the block must really be grouped by this rank's coarse bins, in order
        world, rank = dist.get_world_size(), dist.get_rank()
        bins = np_owner_of(k, world * len(bin_counts)) - rank * len(bin_counts)
        assert (np.diff(bins) >= 0).all() and (np.bincount(bins, minlength=len(bin_counts)) == np.asarray(bin_counts)).all()
        self._add(k, np.ones(len(k), dtype=np.uint64))

    def insert(self, keys, counts=None):
        k = keys.numpy().view(np.uint64)
        self._add(k, np.ones(len(k), dtype=np.uint64) if counts is None else counts.numpy().view(np.uint64))

    def finalize(self, want_summary=True):
        return dict(n_windows=int(self.counts.sum()), n_distinct=len(self.keys), n_records=self.n_records, n_bases=self.n_bases,
                    max_count=int(self.counts.max()) if len(self.counts) else 0)

    def export(self, min_count=1):
        m = self.counts >= min_count
        return self.keys[m], self.counts[m]

    def histogram(self, min_count=1):
        return orc.histogram(self.counts, min_count)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, k, q, data_path, out_dir, n_chunks=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with open(data_path, "rb") as f:
            seq_np, qual_np, offsets_np = pickle.load(f)
        total = len(seq_np)
        a, b = slice_for_rank(total, world, rank, k)
        inside = [int(o) - a for o in offsets_np[1:-1] if a < int(o) < b]
        off = torch.tensor([0] + inside + [b - a], dtype=torch.int64)
        seq = torch.from_numpy(seq_np[a:b].copy())
        qual = None if qual_np is None else torch.from_numpy(qual_np[a:b].copy())
        sc = ShardedKmerCounter(OracleShardEngine(k, q), n_chunks=n_chunks)
        sc.count(seq, off, qual)
        summary = sc.finalize()
        keys, counts = sc.export_gathered(1)
        hv, hf = sc.histogram(2)
        lk, _ = sc.engine.export(1)
        assert (np_owner_of(lk, world) == rank).all()  # every key lives on its owner shard only
        with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
            pickle.dump(dict(summary={x: summary[x] for x in ("n_windows", "n_distinct", "n_records", "max_count")},
                             keys=keys, counts=counts, hv=hv, hf=hf, sent=sc.sent_keys, recv=sc.recv_keys), f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k,q,n_chunks", [(2, 21, None, 0), (2, 5, 20, 0), (3, 31, 10, 0), (2, 21, 15, 3), (3, 7, None, 5)])
def test_sharded_count_equals_single_process(tmp_path, world, k, q, n_chunks):
    rng = np.random.default_rng(world * 100 + k)
    recs, quals = [], []
    for _ in range(40):
        n = int(rng.integers(0, 700))
        recs.append(bytes(rng.choice(list(b"ACGT" * 10 + b"N"), size=n).tolist()))
        quals.append(bytes((rng.integers(0, 42, size=n) + 33).astype(np.uint8).tolist()))
    seq, qual, offsets = orc.make_batch(recs, quals)
    if q is None:
        qual = None
    data = tmp_path / "data.pkl"
    with open(data, "wb") as f:
        pickle.dump((seq, qual, offsets), f)
    mp.spawn(_worker, args=(world, _free_port(), k, q, str(data), str(tmp_path), n_chunks), nprocs=world, join=True)
    okeys, ocounts, windows = orc.count_batch(k, seq, qual, offsets, q, mode="literal")
    ov, of = orc.histogram(ocounts, 2)
    sent = recv = 0
    for r in range(world):
        with open(tmp_path / f"r{r}.pkl", "rb") as f:
            res = pickle.load(f)
        assert (res["keys"] == okeys).all() and (res["counts"] == ocounts).all()
        assert res["summary"]["n_windows"] == windows and res["summary"]["n_distinct"] == len(okeys)
        assert res["summary"]["max_count"] == (int(ocounts.max()) if len(ocounts) else 0)
        assert (res["hv"] == ov).all() and (res["hf"] == of).all()
        sent += res["sent"]; recv += res["recv"]
    assert sent == recv  # every key sent is received exactly once


def test_slices_partition_the_windows():
    """Each window belongs to exactly one rank's slice (k-1 halo, no double counting)."""
    for total, world, k in ((1000, 2, 21), (1000, 8, 32), (97, 4, 5), (50, 8, 31)):
        owners = np.zeros(max(0, total - k + 1), dtype=np.int64)
        for r in range(world):
            a, b = slice_for_rank(total, world, r, k)
            for s in range(a, max(a, b - k + 1)):
                owners[s] += 1
        assert (owners == 1).all()


def test_owner_function_matches_library_and_merge():
    keys = np.random.default_rng(1).integers(0, 2**62, 500, dtype=np.uint64)
    for n in (2, 3, 8, 4096):
        assert np_owner_of(keys, n).tolist() == [kb.owner_of(int(x), n) for x in keys]
    v, f = merge_histograms([(np.array([1, 3], dtype=np.uint64), np.array([2, 1], dtype=np.uint64)),
                             (np.array([1, 2], dtype=np.uint64), np.array([5, 7], dtype=np.uint64))])
    assert v.tolist() == [1, 2, 3] and f.tolist() == [7, 7, 1]
