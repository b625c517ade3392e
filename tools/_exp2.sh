#!/bin/bash
# N=2 exchange A/B (run with gpurun --gpus 2)
mkdir -p gpurun_out
for ex in peer host nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex --delta-steps 100 --batch 256 > gpurun_out/bench_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err
  echo "exchange $ex rc=$?"
  tail -2 gpurun_out/bench_n2_$ex.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n2_$ex.json'))
    print('$ex', 'value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e_ms', d['e2e']['ms_per_step'], 'prep_us', d['e2e']['host_prepare_us'], 'inc_e2e_ms', d['incremental']['e2e_ms_per_eval'], 'inc_dev_ms', d['incremental']['device_ms_per_eval'], 'inc_prep', d['incremental']['host_prepare_us_per_eval'], 'batch_ms', d['batch']['ms_per_batch'], 'prob', d['result'])
except Exception as e:
    print('$ex failed', e)
PY
done
