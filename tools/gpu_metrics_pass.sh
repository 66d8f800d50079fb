#!/bin/bash
# one ncu metrics pass over every kernel of one time step (4 M column): time, instructions, issue utilisation, DRAM bytes
out=${1:-gpurun_out/metrics}
mkdir -p $out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
for mode in tile list; do
  if [ $mode = list ]; then export SPSPH_TILE=0; else export SPSPH_TILE=1; fi
  ncu --metrics $M --clock-control none -s ${SKIP:-70} -c ${COUNT:-48} --csv --log-file $out/metrics_$mode.csv \
      python tools/run_steps.py --steps 3 > $out/run_$mode.log 2>&1
  unset SPSPH_TILE
done
ls -la $out
