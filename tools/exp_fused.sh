#!/bin/bash
# A/B of the fused-reservation rows kernels (KMG_FUSED) on a GPU box: parity subset, then the C4 bench both ways.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
KMG_FUSED=1 timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c4_bin_geometry or sharded_exchange_shapes or input_outgrows" > gpurun_out/exp1_pytest_fused1.log 2>&1
echo "pytest fused=1 rc=$?"; tail -3 gpurun_out/exp1_pytest_fused1.log
KMG_FUSED=0 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c4_bin_geometry_on_skewed" > gpurun_out/exp1_pytest_fused0.log 2>&1
echo "pytest fused=0 rc=$?"; tail -3 gpurun_out/exp1_pytest_fused0.log
for f in 0 1 0 1; do
  KMG_FUSED=$f timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/exp1_bench_fused$f.json 2>> gpurun_out/exp1_bench_fused$f.err
  echo "bench fused=$f rc=$?"
  tail -1 gpurun_out/exp1_bench_fused$f.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['pipeline'])"
done
