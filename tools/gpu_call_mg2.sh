#!/bin/bash
# 2-GPU call of the closing build: the slab tests on real NCCL (incl. bench.py --gpus 2 inside pytest)
out=gpurun_out/mg2
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $out/multi_gpu.log 2>&1; echo "multi-GPU tests: exit $?"; tail -5 $out/multi_gpu.log
