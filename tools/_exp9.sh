#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "patch or c4_shape or there_and_back or capacity or golden" 2>&1 | tail -15 > gpurun_out/patch_tests.log
echo "tests rc=$?"; tail -5 gpurun_out/patch_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/p_c2.json 2> gpurun_out/p_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload c4shard > gpurun_out/p_c4.json 2> gpurun_out/p_c4.err; echo "c4 rc=$?"
tail -3 gpurun_out/p_c4.err
python - <<'PY'
import json
for f in ('p_c2','p_c4'):
    try: d=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as e: print(f, 'no json', e); continue
    for k in ('value','ms_per_step','e2e','e2e_same_walks','e2e_cold_list','sa_iters_per_s','incremental'):
        v=d.get(k)
        if isinstance(v,dict): v={kk:vv for kk,vv in v.items() if kk not in ('note','mix')}
        print(f, k, str(v)[:400])
    print(f, 'frac', d['roofline']['frac'])
PY
