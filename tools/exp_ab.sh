#!/bin/bash
# A/B of a kernel variant switched by an environment variable on a GPU box: parity subset with the variant on, then the C4 bench both ways.
#   tools/exp_ab.sh KMG_ROWS2 [pytest -k expression]
VAR=${1:-KMG_ROWS2}
KEXPR=${2:-"c4_bin_geometry or sharded_exchange_shapes or input_outgrows or skewed_high or phase_b_sieve"}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
env $VAR=1 timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$KEXPR" > gpurun_out/ab_${VAR}_pytest.log 2>&1
echo "pytest $VAR=1 rc=$?"; tail -3 gpurun_out/ab_${VAR}_pytest.log
for f in 0 1 0 1; do
  env $VAR=$f timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/ab_${VAR}_$f.json 2>> gpurun_out/ab_${VAR}_$f.err
  echo "bench $VAR=$f rc=$?"
  tail -1 gpurun_out/ab_${VAR}_$f.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['pipeline']; print(round(d['ms_per_step'],2), 'a1', round(p['a1_ms'],2), 'a2', round(p['a2_ms'],2), 'b', round(p['phase_b_ms'],2))"
done
