#!/bin/bash
# ncu artefacts of the round (1 GPU): launch list of a short bench run, full capture of the streaming kernel on both workloads
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --delta-steps 6 --batch 64 --no-dropin"
$CMD > gpurun_out/plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_raw.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:paired_stream -s 6 -c 1 -o gpurun_out/r02_stream_c2 -f $CMD > gpurun_out/ncu2.log 2>&1; echo "full c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:paired_stream -s 6 -c 1 -o gpurun_out/r02_stream_c4 -f $CMD --workload c4shard > gpurun_out/ncu3.log 2>&1; echo "full c4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:paired_delta -s 3 -c 1 -o gpurun_out/r02_delta_c2 -f $CMD > gpurun_out/ncu4.log 2>&1; echo "full delta rc=$?"
ls -la gpurun_out/*.ncu-rep
