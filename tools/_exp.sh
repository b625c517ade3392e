timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/v_pytest.txt
for v in 0 2; do
  GAML_STREAM_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --delta-steps 10 --batch 0 > gpurun_out/v_bench_$v.json 2> gpurun_out/v_bench_$v.err
done
for v in 0; do
GAML_STREAM_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --delta-steps 10 --batch 0 --workload c4shard > gpurun_out/v_c4_$v.json 2> gpurun_out/v_c4_$v.err
done
