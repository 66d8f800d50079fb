#!/bin/bash
# ncu --set full captures of list-path kernels: KERNELS="name:skip ..." (at most ~5 per call: 64 MiB limit of gpurun_out)
out=${1:-gpurun_out/ncu_list}
mkdir -p $out
for ks in ${KERNELS:-k_sweep_a_sp:7 k_sweep_b_sp:7 k_artvisc:7 k_fill:1 k_count:1}; do
  k=${ks%%:*}; s=${ks##*:}
  ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s $s -c 1 -f -o $out/$k \
      python tools/run_steps.py --steps 2 > $out/$k.log 2>&1
done
ls $out
