#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --delta-steps 200 --batch 1024 --no-dropin"
ncu --set full --clock-control none --import-source on -k regex:batch_touch -s 1 -c 1 -o gpurun_out/r02_batch_touch_c2 -f $CMD > gpurun_out/ncu5.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r02_batch_touch_c2.ncu-rep
