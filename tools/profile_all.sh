#!/bin/bash
# ncu evidence for profiles/: launch list of one time step + one `--set full` capture per hot kernel
# (4 M-particle refined Bui column, 1 GPU). Run on the GPU box: bash tools/profile_all.sh [outdir]
out=${1:-gpurun_out/prof}
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 49 -c 60 --csv --log-file $out/launches.csv \
    python tools/run_steps.py --steps 2 > $out/launch_run.log 2>&1
for k in k_fill k_count k_move; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}$" -s 1 -c 1 -f -o $out/$k \
      python tools/run_steps.py --steps 2 > $out/$k.log 2>&1
done
for k in k_sweep_a_sp k_sweep_a_node k_sweep_b_node k_sweep_b_sp k_artvisc; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s 7 -c 1 -f -o $out/$k \
      python tools/run_steps.py --steps 2 > $out/$k.log 2>&1
done
python tools/run_steps.py --warmup 3 --steps 10 --profile > $out/profile.log 2>&1
ls -la $out
