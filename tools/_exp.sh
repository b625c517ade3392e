timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/v_pytest.txt
timeout 900 python bench.py > gpurun_out/v_full.json 2> gpurun_out/v_full.err
timeout 300 python bench.py --workload c4shard --no-cpu-baseline --batch 0 --delta-steps 20 > gpurun_out/v_c4.json 2> gpurun_out/v_c4.err
