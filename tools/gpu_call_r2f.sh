#!/bin/bash
out=gpurun_out/r2f
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > $out/gpu_suite.log 2>&1
echo "gpu suite: exit $?"; tail -5 $out/gpu_suite.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
tail -c 3000 $out/bench.json; tail -5 $out/bench.err
