#!/bin/bash
# round 2, first GPU call: GPU suite on the new tile path, per-kernel event profile of the 4 M column on both paths,
# bench line, launch list. Everything lands in gpurun_out/r2a/.
out=gpurun_out/r2a
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $out/gpu_suite.log 2>&1
echo "gpu suite: exit $?"; tail -5 $out/gpu_suite.log
timeout 600 python tools/run_steps.py --warmup 3 --steps 10 --profile > $out/profile_tile.log 2>&1
cat $out/profile_tile.log
SPSPH_TILE=0 timeout 600 python tools/run_steps.py --warmup 3 --steps 10 --profile > $out/profile_list.log 2>&1
head -1 $out/profile_list.log
timeout 900 python bench.py --no-cpu > $out/bench.json 2> $out/bench.err
tail -c 2500 $out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $out/launches.csv \
    python tools/run_steps.py --steps 3 > $out/launch_run.log 2>&1
tail -2 $out/launch_run.log
