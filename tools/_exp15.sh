#!/bin/bash
mkdir -p gpurun_out
for wlk in c4shard; do
for t in 0 1 2; do
  GAML_B200_T2_FIRST=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 0 --no-other-configs --no-dropin --delta-steps 8 --workload $wlk > gpurun_out/t2_${wlk}_$t.json 2> gpurun_out/t2_${wlk}_$t.err
  python - <<PY
import json
d=json.load(open('gpurun_out/t2_${wlk}_$t.json'))
r=d['roofline']
print('$wlk t2_first=$t value us', round(d['ms_per_step']*1e3,2), 'kernel us', round(r['kernel_ms']*1e3,2), 'min', round(r['kernel_ms_min']*1e3,2), 'frac', round(r['frac'],3), 'e2e us', round(d['e2e']['ms_per_step']*1e3,1), 'timeline', d.get('device_timeline_us'), 'prob', d['result']['prob'])
PY
done
done
