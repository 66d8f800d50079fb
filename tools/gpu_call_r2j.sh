#!/bin/bash
out=gpurun_out/r2j
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_full_size.py -m gpu -x -q > $out/gpu_parity.log 2>&1
echo "parity: exit $?"; tail -3 $out/gpu_parity.log
timeout 600 python tools/run_steps.py --warmup 3 --steps 10 --profile > $out/profile_list.log 2>&1
cat $out/profile_list.log
