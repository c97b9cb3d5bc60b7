#!/bin/bash
# full GPU validation on a box: whole -m gpu suite, default bench line (C4), launch list of one step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/final_pytest.log
timeout 600 python bench.py > gpurun_out/final_bench_C4.json 2> gpurun_out/final_bench_C4.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_C4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'frac_job', d['roofline']['frac_job'], d['roofline']['pipeline'], 'file', d.get('e2e_file',{}).get('value'))
PY
