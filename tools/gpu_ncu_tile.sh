#!/bin/bash
# ncu --set full captures of the tile kernels (4 M column); outputs in gpurun_out/ncu_tile/
out=${1:-gpurun_out/ncu_tile}
mkdir -p $out
for k in ${KERNELS:-k_tile_b_sp k_tile_a_sp k_tile_b_node k_tile_move k_tile_build}; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s ${SKIP:-2} -c 1 -f -o $out/$k \
      python tools/run_steps.py --steps 2 > $out/$k.log 2>&1
done
ls -la $out
