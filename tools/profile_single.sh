set -x
mkdir -p gpurun_out/r1b
for k in k_fill k_count k_move k_rank; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}$" -s 1 -c 1 -f -o gpurun_out/r1b/$k python tools/run_steps.py --steps 2 > gpurun_out/r1b/$k.log 2>&1
done
for k in "k_sweep_a_sp" "k_sweep_a_node" "k_sweep_b_node" "k_sweep_b_sp"; do
  ncu --set full --clock-control none --import-source on -k "regex:^${k}" -s 7 -c 1 -f -o gpurun_out/r1b/${k}_s7 python tools/run_steps.py --steps 2 > gpurun_out/r1b/${k}_s7.log 2>&1
done
ls -la gpurun_out/r1b
