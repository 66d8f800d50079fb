#!/bin/bash
# 2-GPU call: the slab tests (incl. bench.py --gpus 2 inside pytest) and one bench line with rank profiles
out=gpurun_out/mg
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 1800 python -m pytest tests/test_multi_gpu.py -m gpu -q > $out/multi_gpu.log 2>&1; echo "multi-GPU tests: exit $?"; tail -5 $out/multi_gpu.log
bash tools/gpu_scale.sh 2 $out > $out/scale2.log 2>&1; tail -c 600 $out/bench_n2.json; tail -3 $out/bench_n2.err
