#!/bin/bash
# round-2 artefacts (1 GPU): the default bench line, launch list of a short run, full captures of the streaming kernel on both workloads
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_n1_reference.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 --workload c4shard --no-cpu-baseline > gpurun_out/r02_bench_n1_c4shard.json 2> gpurun_out/r02_bench_n1_c4shard.err; echo "bench c4 rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --delta-steps 6 --batch 64 --no-dropin --no-other-configs"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_raw.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:paired_stream -s 6 -c 1 -o gpurun_out/r02_stream_c2 -f $CMD > gpurun_out/ncu2.log 2>&1; echo "full c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:paired_stream -s 6 -c 1 -o gpurun_out/r02_stream_c4 -f $CMD --workload c4shard > gpurun_out/ncu3.log 2>&1; echo "full c4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:paired_delta -s 3 -c 1 -o gpurun_out/r02_delta_c2 -f $CMD > gpurun_out/ncu4.log 2>&1; echo "full delta rc=$?"
