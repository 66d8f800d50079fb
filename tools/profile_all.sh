set -x
mkdir -p gpurun_out/r1b
ncu --metrics gpu__time_duration.sum --clock-control none -s 46 -c 60 --csv --log-file gpurun_out/r1b/launches.csv python tools/run_steps.py --steps 2 > gpurun_out/r1b/launch_run.log 2>&1
for k in "k_sweep_b_sp<0>" "k_sweep_b_node<0>" "k_sweep_a_sp<0, 0>" "k_sweep_a_node<0, 0, 0>" k_fill k_count k_artvisc k_move; do
  n=$(echo "$k" | tr -d '<>, ')
  ncu --set full --clock-control none --import-source on -k "regex:^${k%%<*}" -s 5 -c 1 -f -o gpurun_out/r1b/$n python tools/run_steps.py --steps 2 > gpurun_out/r1b/$n.log 2>&1
done
python tools/run_steps.py --steps 10 --profile > gpurun_out/r1b/profile.log 2>&1
ls -la gpurun_out/r1b
