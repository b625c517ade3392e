#!/bin/bash
mkdir -p gpurun_out
timeout 900 bash tools/sa_rate.sh gpurun_out/sa_rate 4000 100 400000 200 > gpurun_out/sa_rate.log 2>&1; echo "sa_rate rc=$?"
tail -8 gpurun_out/sa_rate.log
