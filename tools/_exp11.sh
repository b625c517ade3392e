#!/bin/bash
GAML_B200_PREP_TIMING=1 timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload c4shard --no-dropin --batch 0 > gpurun_out/t_c4.json 2> gpurun_out/t_c4.err; echo "c4 rc=$?"
grep "prepare us" gpurun_out/t_c4.err | head -20
