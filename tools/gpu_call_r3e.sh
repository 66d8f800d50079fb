bash tools/variants.sh run gpurun_out/var_r3e base fill7 fill6 fill5
echo "--- single stream"
SPSPH_DUAL_STREAM=0 python tools/run_steps.py --deck /tmp/spsph_variant_deck --warmup 3 --steps 10 --profile | grep -E "ms/step,|k_sweep"
