#!/usr/bin/env python
"""Headline benchmark: canonical k-mers counted per second on the 3.1 Gbp k=21 workload
(BASELINE.json configs[3], "C4"), hash-sharded across N B200s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU algorithm (oracle port)

A "step" is one complete counting job: clear the table, scan the synthetic genome (ASCII already in
HBM for `value`; pinned HOST memory for `e2e`), bucket/exchange/upsert (N>1), table final.
Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "canonical k-mers counted/sec"
UNIT = "kmers/s"
SEED = 44                       # G3100
B_ALG_HASH_NEW = 0.375 + 32.0   # algorithmic bytes per counted k-mer, every key new


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--bases", type=float, default=3.1e9, help="total bases of the synthetic genome")
    ap.add_argument("--records", type=int, default=31)
    ap.add_argument("--cpu-sample-bases", type=float, default=32e6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        # "under load" = the upper half of the samples (idle samples before/after the region drag the median down)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_slice(total: int, records: int, world: int, rank: int, k: int):
    """Global stream = `records` equal records back to back.  Returns this rank's byte range and the record-start
    offsets inside it (relative), per krust_b200.dist.slice_for_rank."""
    import numpy as np
    from krust_b200.dist import slice_for_rank
    a, b = slice_for_rank(total, world, rank, k)
    rec_len = total // records
    starts = [r * rec_len for r in range(records)]
    inside = [s - a for s in starts if a < s < b]
    offsets = np.array([0] + inside + [b - a], dtype=np.uint64)
    return a, b, offsets


def expected_windows(total: int, records: int, k: int) -> int:
    rec_len = total // records
    last = total - rec_len * (records - 1)
    return (records - 1) * max(0, rec_len - k + 1) + max(0, last - k + 1)


def run_reference(args, rank: int):
    """The reference's own CPU algorithm (oracle port: the Rust crate cannot be built here) on a bounded
    sample of the same workload, all host threads."""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    n = int(args.cpu_sample_bases)
    rec = max(1, args.records)
    rec_len = n // rec
    n = rec_len * rec
    seq = orc.synth_uniform(SEED, 0, n)
    offsets = np.arange(0, n + 1, rec_len, dtype=np.uint64)
    for _ in range(args.warmup):
        orc.reference_path_count(args.k, seq[: n // 8], None, (offsets // 8).astype(np.uint64), threads=cores)
    t0 = time.perf_counter()
    windows = 0
    for _ in range(args.steps):
        w, _d = orc.reference_path_count(args.k, seq, None, offsets, threads=cores)
        windows += w
    dt = time.perf_counter() - t0
    value = windows / dt
    sample = f"{rec} records x {rec_len} bp uniform ACGT (seed {SEED}), k={args.k}, per step; restated reference CPU path (oracle/kmer_oracle.c orc_reference_path_count)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "C4: k=21 canonical k-mer counting, 3.1 Gbp uniform-random FASTA (31 x 100 Mbp), bounded CPU sample",
                       "k": args.k, "sample_bases": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """The driver parses stdout as ONE JSON line: route everything libraries print there (e.g. NCCL's version banner)
    to stderr and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    claim_stdout()
    if os.environ.get("KMG_BENCH_DEBUG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["KMG_BENCH_DEBUG"]), repeat=True, file=sys.stderr)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import krust_b200 as kb
    from krust_b200 import _lib
    from krust_b200.dist import GpuShardEngine, ShardedKmerCounter

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: krust_b200 has no CPU fallback")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()

    k = args.k
    total = int(args.bases)
    a, b, offsets_np = workload_slice(total, args.records, world, rank, k)
    n_local = b - a
    exp_windows = expected_windows(total, args.records, k)

    # ---- synthetic input, generated on the device (same counter-based generator as the oracle)
    d_seq = torch.empty(n_local + 64, dtype=torch.uint8, device=dev)[:n_local]
    d_off = torch.from_numpy(offsets_np.astype(np.int64)).to(dev)
    engine = GpuShardEngine(k, dev, expected_distinct=int(exp_windows / world * 1.03) + 1024, flags=args.flags)
    engine.counter.synth_uniform_device(SEED, a, n_local, d_seq.data_ptr())
    sharded = ShardedKmerCounter(engine)
    off_arg = d_off if len(offsets_np) > 2 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_device():
        t_dbg = time.perf_counter()
        engine.reset()
        if os.environ.get("KMG_DIST_TIMING") and rank == 0:
            torch.cuda.synchronize(dev); print(f"[dist timing] reset: {(time.perf_counter() - t_dbg) * 1e3:.2f} ms", file=sys.stderr, flush=True); t_dbg = time.perf_counter()
        sharded.count(d_seq, off_arg, expected_keys_per_rank=exp_windows // world + 1024)
        if os.environ.get("KMG_DIST_TIMING") and rank == 0:
            t_dbg = time.perf_counter()
        engine.finalize(False)
        if os.environ.get("KMG_DIST_TIMING") and rank == 0:
            torch.cuda.synchronize(dev); print(f"[dist timing] finalize: {(time.perf_counter() - t_dbg) * 1e3:.2f} ms", file=sys.stderr, flush=True)

    # ---- device-resident timing: W warm-up, then exactly K steps between barriers, CUDA events, max over ranks
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.kmg_kernel_launches()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    launches = L.kmg_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    summary = sharded.finalize()   # global sums over shards
    if summary["n_windows"] != exp_windows:
        raise SystemExit(f"PARITY FAILURE: counted {summary['n_windows']} windows, expected {exp_windows}")
    local = summary.get("local", summary)
    value = exp_windows / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel, timed live with CUDA events on its launching stream
    peak, peak_src = peaks()
    pipeline = None
    gpu_windows = exp_windows / world   # per-GPU share (strong scaling)
    if local["path"] == 2:
        # partitioned pipeline: phase A (A1 scan + coarse scatter, A2 refine) and phase B (count_partitions kernel); reset()
        # clears the timers, so these are the last step's kernels on rank 0.  Algorithmic bytes per k-mer: A1 = 0.375 in + 8 out,
        # A2 = 8 in + 8 out, B = 8 in + 16 out; design-independent figure for the whole job: 32.375.
        a_ms, b_ms = local["scan_ns"] / 1e6, local["consolidate_ns"] / 1e6
        pipeline = {"phase_a_ms": a_ms, "phase_b_ms": b_ms, "phase_a_gbs": gpu_windows * 24.375 / (a_ms * 1e-3) / 1e9 if a_ms else 0.0,
                    "phase_b_gbs": gpu_windows * 24.0 / (b_ms * 1e-3) / 1e9 if b_ms else 0.0,
                    "whole_gbs": gpu_windows * B_ALG_HASH_NEW / ((a_ms + b_ms) * 1e-3) / 1e9 if a_ms + b_ms else 0.0}
        # dominant single KERNEL: phase B is one launch; phase A is five (ingest, partition_count, partition_scatter_staged,
        # refine<count>, refine<scatter>), the largest of which takes ~40 % of phase A (profiles/r1_v5_launches.csv)
        if b_ms >= 0.4 * a_ms:
            kern_ms, per_unit, kern_name = b_ms, 24.0, "count_partitions_smem_kernel (phase B: one CTA per hash partition, upsert into a shared-memory table, compact)"
        else:
            kern_ms, per_unit, kern_name = a_ms, 24.375, "phase A kernels (partition_count/scatter over the scan + refine count/scatter: two-level hash partitioning)"
        alg_bytes = gpu_windows * per_unit
    else:
        kern_ms = local["kernel_ns"] / 1e6          # scan_count_kernel of the last step (reset() clears the timer)
        alg_bytes = gpu_windows * B_ALG_HASH_NEW
        kern_name = "scan_count_kernel (tile scan + upsert into one HBM-resident table / direct-indexed array)"
        per_unit = B_ALG_HASH_NEW
    # DRAM traffic of the dominant kernel: bytes per k-mer from the committed `ncu --set full` capture (taken on a 5e8-base
    # slice of the same workload: the full job would make ncu save/restore ~100 GB per pass), scaled to this launch
    traffic = None
    try:
        tpk = json.load(open(os.path.join(ROOT, "profiles", "traffic_per_kmer.json")))
        for name, v in tpk.items():
            if name in kern_name:
                traffic = v * gpu_windows
    except Exception:
        traffic = None
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kern_name, "kernel_ms": kern_ms, "alg_bytes_per_kmer": per_unit, "peak_source": peak_src,
                "pipeline": pipeline}

    # ---- end to end: HOST (pinned) ASCII in, histogram + summary out, through the public C-ABI call
    e2e = None
    if not args.no_e2e:
        h_seq = torch.empty(n_local, dtype=torch.uint8, pin_memory=True)
        h_seq.copy_(d_seq)
        torch.cuda.synchronize(dev)
        h_np = h_seq.numpy()
        d2h_bytes = 0

        def step_e2e():
            nonlocal d2h_bytes
            engine.reset()
            if world == 1:
                engine.counter.count_batch(h_np, None, offsets_np)     # kmg_count_ascii: chunked H2D overlapped with the kernels
            else:
                d_tmp = h_seq.to(dev, non_blocking=True)
                sharded.count(d_tmp, off_arg, expected_keys_per_rank=exp_windows // world + 1024)
            engine.finalize(False)
            vals, freqs = sharded.histogram(1)                          # D2H of the result
            d2h_bytes = (vals.nbytes + freqs.nbytes) + 65536 * 8
            return vals, freqs

        for _ in range(max(2, args.warmup // 2)):   # the first host-fed step sizes the staging ring and the pool
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            vals, freqs = step_e2e()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        if int((vals * freqs).sum()) != exp_windows:
            raise SystemExit("PARITY FAILURE in the end-to-end path")
        e2e = {"value": exp_windows / (dt / args.steps), "unit": UNIT, "h2d_bytes_per_step": int(n_local + offsets_np.nbytes),
               "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": dt / args.steps * 1e3,
               "result": "count-of-counts histogram + summary to host; table stays in HBM"}
        del h_seq

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only; bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        cores = os.cpu_count() or 1
        rec_len = int(args.cpu_sample_bases) // args.records
        n = rec_len * args.records
        seq = orc.synth_uniform(SEED, 0, n)
        offs = np.arange(0, n + 1, rec_len, dtype=np.uint64)
        t0 = time.perf_counter()
        w, _d = orc.reference_path_count(k, seq, None, offs, threads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": w / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.records} records x {rec_len} bp of the same generator (seed {SEED}), k={k}; restated reference CPU path, {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic",
                "config": {"workload": "C4: k=21 canonical k-mer counting, 3.1 Gbp uniform-random FASTA (31 records x 100 Mbp), "
                                       "hash-sharded across N GPUs",
                           "k": k, "bases": total, "records": args.records, "windows": exp_windows,
                           "distinct": summary["n_distinct"], "path": {0: "hbm-table", 1: "direct-4^k", 2: "partitioned"}[local["path"]],
                           "table_slots_or_partitions_per_gpu": local["table_capacity"],
                           "l2_policy": "inputs and table are far larger than L2 (no flush needed)",
                           "step": "table clear + ingest + scan/upsert (+ bucket, all-to-all, upsert for N>1) + finalize",
                           "parallelism": f"hash-shard x{world}" if world > 1 else "single GPU"},
                "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches)}
        emit(line)
    engine.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
