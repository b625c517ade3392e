#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "append or grows or annealing_driver or dropin or flat_cache" 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 0 --no-other-configs > gpurun_out/a_c2.json 2> gpurun_out/a_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 0 --workload c4shard > gpurun_out/a_c4.json 2> gpurun_out/a_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ('a_c2','a_c4'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:v for k,v in d.get('append',{}).items() if k!='note'})
    print(f, {k:v for k,v in d.get('dropin',{}).items() if k not in ('note','what_is_compared')})
PY
