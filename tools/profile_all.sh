#!/bin/bash
# ncu evidence for profiles/ (4 M-particle refined Bui column, 1 GPU). A gpurun call brings back at most 64 MiB, one
# --set full capture is ~10 MB, so the captures are split over two calls:
#   gpurun -- 'bash tools/profile_all.sh A'   launch list + event profile + per-kernel metric tables (both paths)
#                                             + full captures of the five pair-sum kernels of the id-list path
#   gpurun -- 'bash tools/profile_all.sh B'   full captures of the neighbour build / position update (id-list path)
#                                             and of the cell-tile path's build and gradient sweep
# then here: python tools/make_profile_summaries.py gpurun_out/prof_A ; python tools/make_profile_summaries.py gpurun_out/prof_B
part=${1:-A}
out=gpurun_out/prof_$part
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
cap() {  # kernel regex, launches to skip, output name, extra env
  env $4 ncu --set full --clock-control none --import-source on -k "regex:^$1" -s $2 -c 1 -f -o $out/$3 \
      python tools/run_steps.py --steps 2 > $out/$3.log 2>&1
}
if [ $part = A ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 49 -c 60 --csv --log-file $out/launches.csv \
      python tools/run_steps.py --steps 2 > $out/launch_run.log 2>&1
  python tools/run_steps.py --warmup 3 --steps 10 --profile > $out/profile.log 2>&1
  bash tools/gpu_metrics_pass.sh $out > /dev/null 2>&1
  for k in k_sweep_a_sp k_sweep_a_node k_sweep_b_node k_sweep_b_sp k_artvisc; do cap $k 7 $k SPSPH_TILE=0; done
else
  for k in k_fill k_count k_move; do cap $k 1 $k SPSPH_TILE=0; done
  cap k_tile_build 4 k_tile_build SPSPH_TILE=1   # the stress-particle build of step 2 (launch order: <0>, <1>, <2> per step)
  cap k_tile_b_sp 2 k_tile_b_sp SPSPH_TILE=1
fi
ls -la $out
