#!/bin/bash
# 2-GPU call: remaining single-GPU tests, the multi-GPU tests, bench at N = 2
out=gpurun_out/r2g
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_tile_path.py tests/test_multi_gpu.py tests/test_zz_gpu_new_paths.py -m gpu -q > $out/gpu_tests.log 2>&1
echo "tests: exit $?"; tail -8 $out/gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --no-cpu > $out/bench_n2.json 2> $out/bench_n2.err
tail -c 2500 $out/bench_n2.json; tail -3 $out/bench_n2.err
