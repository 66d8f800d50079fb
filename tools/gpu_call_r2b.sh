#!/bin/bash
out=gpurun_out/r2c
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q > $out/gpu_parity.log 2>&1
echo "parity: exit $?"; tail -3 $out/gpu_parity.log
timeout 600 python tools/run_steps.py --warmup 3 --steps 10 --profile > $out/profile_tile.log 2>&1
cat $out/profile_tile.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file $out/launches.csv \
    python tools/run_steps.py --steps 3 > $out/launch_run.log 2>&1
KERNELS="k_tile_b_sp k_tile_a_sp" SKIP=2 bash tools/gpu_ncu_tile.sh $out > /dev/null 2>&1
KERNELS="k_tile_build k_tile_move" SKIP=1 bash tools/gpu_ncu_tile.sh $out > /dev/null 2>&1
ls $out
