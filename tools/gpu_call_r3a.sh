#!/bin/bash
# round 2, session 2, call a: lean k_count + threaded upload analysis on hardware (parity, per-kernel times, bench line)
out=gpurun_out/r3a
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -x -q > $out/parity.log 2>&1; echo "parity exit $?"; tail -3 $out/parity.log
timeout 600 python tools/run_steps.py --deck /tmp/spsph_deck --warmup 3 --steps 10 --profile > $out/profile_list.log 2>&1; cat $out/profile_list.log
timeout 900 python bench.py --no-extras --no-cpu > $out/bench.log 2>&1; echo "bench exit $?"; tail -1 $out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print(d['roofline']['kernels_ms_per_step'])"
