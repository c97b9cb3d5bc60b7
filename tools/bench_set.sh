#!/bin/bash
# bench lines of the other configurations + the ncu launch list of one C4 step (profiles/)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for c in C1 C2 C3; do
  timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_r4_$c.json 2> gpurun_out/bench_r4_$c.err; echo "$c rc=$?"
  tail -1 gpurun_out/bench_r4_$c.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value']/1e9, 'e2e', d['e2e']['value']/1e9 if d.get('e2e') else None, d['e2e'].get('ms_per_step') if d.get('e2e') else None, 'frac', d['roofline']['frac'], d['roofline']['frac_job'], d.get('also'), (d.get('e2e_file') or {}).get('ms_per_step'))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r4_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r4_launches_bench.log 2>&1; echo "ncu rc=$?"
