out=gpurun_out/ncu_r3o; mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
for v in sub2m6 m3; do
  SPSPH_CUDA_SO=$PWD/stress-particle-sph_b200/variants/libspsph_cuda_$v.so ncu --set full --clock-control none --import-source on -k "regex:^k_sweep_b_sp" -s 7 -c 1 -f -o $out/b_sp_$v python tools/run_steps.py --steps 2 > $out/$v.log 2>&1
done
ls $out
