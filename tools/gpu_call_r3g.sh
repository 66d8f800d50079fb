bash tools/variants.sh run gpurun_out/var_r3g base f6 f8
KERNELS="k_fill:1" bash tools/gpu_ncu_list.sh gpurun_out/ncu_r3g > /dev/null 2>&1
