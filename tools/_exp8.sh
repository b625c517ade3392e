#!/bin/bash
# N=8 exchange A/B on the whole of config 4 (run with gpurun --gpus 8)
mkdir -p gpurun_out
for ex in peer host nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --exchange $ex --delta-steps 100 --batch 1024 > gpurun_out/bench_n8_$ex.json 2> gpurun_out/bench_n8_$ex.err
  echo "exchange $ex rc=$?"
  grep -v "^\[W\|^W1\|^\*\*\*\|Setting OMP" gpurun_out/bench_n8_$ex.err | tail -3
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n8_$ex.json'))
    print('$ex', 'value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e_ms', d['e2e']['ms_per_step'], 'prep_us', d['e2e']['host_prepare_us'], 'inc_e2e_ms', d['incremental']['e2e_ms_per_eval'], 'inc_dev_ms', d['incremental']['device_ms_per_eval'], 'inc_prep', d['incremental']['host_prepare_us_per_eval'], 'batch_ms', d['batch']['ms_per_batch'], 'prob', d['result'])
except Exception as e:
    print('$ex failed', e)
PY
done
