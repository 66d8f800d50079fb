#!/bin/bash
# bench.py at N ranks of one box: gpurun --gpus N -- 'bash tools/gpu_scale.sh N'
n=${1:-2}
out=${2:-gpurun_out/scale_r3}
mkdir -p $out
python __graft_entry__.py > $out/build_n$n.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $n --no-cpu --rank-profiles $out/ranks > $out/bench_n$n.json 2> $out/bench_n$n.err
tail -c 1800 $out/bench_n$n.json; tail -2 $out/bench_n$n.err
