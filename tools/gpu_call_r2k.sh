#!/bin/bash
out=gpurun_out/r2k
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 1800 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "reslabbing or tile_path" > $out/multi_gpu.log 2>&1
echo "multi-GPU tests: exit $?"; tail -30 $out/multi_gpu.log
