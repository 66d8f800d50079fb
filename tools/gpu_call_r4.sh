#!/bin/bash
# round-2 closing call: whole single-GPU suite, smoke, default bench line, output-frame probe
out=gpurun_out/r4
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/gpu_suite.log 2>&1; echo "gpu suite exit $?"; tail -4 $out/gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $out/smoke.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; tail -2 $out/bench.err
timeout 300 python tools/frame_probe.py > $out/frame_probe.log 2>&1; echo "probe exit $?"; tail -2 $out/frame_probe.log
tail -c 2500 $out/bench.json
