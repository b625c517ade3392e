#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -15 > gpurun_out/gputests.log
echo "tests rc=$?"; tail -12 gpurun_out/gputests.log
