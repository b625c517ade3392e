timeout 300 python tools/host_overhead.py > gpurun_out/host_overhead.txt 2>&1
