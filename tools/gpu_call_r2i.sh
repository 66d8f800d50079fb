#!/bin/bash
out=gpurun_out/r2i
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 2400 python -m pytest tests/test_multi_gpu.py -m gpu -q > $out/multi_gpu.log 2>&1
echo "multi-GPU tests: exit $?"; tail -15 $out/multi_gpu.log
