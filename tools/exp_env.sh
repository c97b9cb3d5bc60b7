#!/bin/bash
# bench C4 (device-resident leg only) under several environment settings:  tools/exp_env.sh "A=1 B=0" "A=1 B=1" ...
# the first setting also runs the partitioned-pipeline parity subset
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
first=1
for setting in "$@"; do
  tag=$(echo "$setting" | tr ' =' '__')
  if [ $first = 1 ] && [ -z "$NO_TESTS" ]; then
    env $setting timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard_group.py -x -q -m gpu -k "c4_bin_geometry or sharded_exchange_shapes or input_outgrows or skewed_high or phase_b_sieve or fused_exchange" > gpurun_out/env_${tag}_pytest.log 2>&1
    echo "pytest [$setting] rc=$?"; tail -2 gpurun_out/env_${tag}_pytest.log
  fi
  first=0
  for rep in 1 2; do
    env $setting timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/env_${tag}.json 2>> gpurun_out/env_${tag}.err
    echo -n "bench [$setting] rc=$? "
    tail -1 gpurun_out/env_${tag}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['pipeline']; print(round(d['ms_per_step'],2), 'a1', round(p['a1_ms'],2), 'a2', round(p['a2_ms'],2), 'b', round(p['phase_b_ms'],2))"
  done
done
