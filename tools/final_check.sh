#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_default.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'dropin', d.get('dropin',{}).get('dropin_us_per_calcprob'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'steps', d['steps'], d['warmup'])
PY
