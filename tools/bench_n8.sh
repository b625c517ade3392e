#!/bin/bash
# the whole of BASELINE config 4 on 8 GPUs, both arms (what the driver runs at N=8); results into gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "n8 rc=$?"
grep -v "^\[W\|^W1\|^\*\*\*\|Setting OMP" gpurun_out/r02_bench_n8.err | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02_bench_n8_reference.json 2>/dev/null
echo "n8 ref rc=$?"
python - <<'PY'
import json
for f in ('r02_bench_n8','r02_bench_n8_reference'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    for k in ('value','ms_per_step','value_incl_exchange','e2e','e2e_same_walks','sa_iters_per_s','incremental','batch','config'):
        v=d.get(k)
        if isinstance(v,dict): v={kk:vv for kk,vv in v.items() if kk not in ('note','mix')}
        print(f, k, str(v)[:330])
    if d.get('roofline'): print(f, 'frac', d['roofline']['frac'], d['roofline']['kernel_ms'])
PY
