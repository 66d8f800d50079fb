#!/bin/bash
# Build-time variants of the pair-sum kernels, timed in ONE gpurun call.
#   here (no GPU):   bash tools/variant_timing.sh build      -> stress-particle-sph_b200/variants/libspsph_cuda_<name>.so
#                                                                (git-ignored, travels with the snapshot) + ptxas logs
#   on the GPU box:  bash tools/variant_timing.sh [outdir]    -> per variant: parity (3 steps of the shipped Bui column
#                                                                against the oracle, bit for bit) and the event profile
#                                                                of 10 steps of the 4 M-particle column
# The variant library is selected with SPSPH_CUDA_SO (spsph/engine.py); the default build is variant "base".
cd "$(dirname "$0")/.."
vdir=stress-particle-sph_b200/variants
VARIANTS=(
  "base:"
  "t64:-DSPSPH_SWEEP_T=64 -DSPSPH_MINB=8"      # same 16 warps per SM, finer tail
  "t32:-DSPSPH_SWEEP_T=32 -DSPSPH_MINB=16"     # one warp per block
  "t64m10:-DSPSPH_SWEEP_T=64 -DSPSPH_MINB=10"  # 20 warps per SM, 102 registers (spills: see the ptxas log)
  "sub2:-DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2"     # two entries in flight per thread instead of four (80-120 registers)
  "pipe2:-DSPSPH_ELL_PIPE=1 -DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2"  # software-pipelined gathers: 2 consumed + 2 in flight
  "pipe2t64:-DSPSPH_ELL_PIPE=1 -DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2 -DSPSPH_SWEEP_T=64 -DSPSPH_MINB=8"
  "pipe2a4:-DSPSPH_ELL_PIPE=1 -DSPSPH_ELL_SUB=2"   # sweep A (little arithmetic per entry) with 4 consumed + 4 in flight
  "sub2m5:-DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2 -DSPSPH_MINB=5"   # 20 warps per SM at 96 registers, 0-28 B of spills
  "sub2m6:-DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2 -DSPSPH_MINB=6"   # 24 warps per SM at 80 registers, 16-124 B of spills
  "pipe2m5:-DSPSPH_ELL_PIPE=1 -DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2 -DSPSPH_MINB=5"  # same, pipelined (32-80 B of spills in sweep B)
  "pipe2ng6:-DSPSPH_ELL_PIPE=1 -DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2 -DSPSPH_A_NG=6"
)
build_one() {
  local name=$1 flags=$2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -shared \
       -Xptxas -v $flags -Iinclude -Istress-particle-sph_b200/csrc -o $vdir/libspsph_cuda_$name.so \
       stress-particle-sph_b200/csrc/spsph_engine.cu -ldl > $vdir/ptxas_$name.log 2>&1 || echo "BUILD FAILED: $name"
}
if [ "$1" = build ]; then
  mkdir -p $vdir
  for v in "${VARIANTS[@]}"; do build_one "${v%%:*}" "${v#*:}" & done
  wait
  ls -la $vdir
  exit 0
fi
out=${1:-gpurun_out/variants}
mkdir -p $out $vdir
python __graft_entry__.py > $out/build.log 2>&1
for v in "${VARIANTS[@]}"; do  # libraries that did not travel with the snapshot: built here, all at once
  [ -f $vdir/libspsph_cuda_${v%%:*}.so ] || build_one "${v%%:*}" "${v#*:}" &
done
wait
for v in "${VARIANTS[@]}"; do
  name=${v%%:*}
  so=$PWD/$vdir/libspsph_cuda_$name.so
  [ -f $so ] || continue
  SPSPH_CUDA_SO=$so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/parity_$name.log 2>&1
  echo "$name parity exit $?" | tee -a $out/summary.txt
  SPSPH_CUDA_SO=$so timeout 600 python tools/run_steps.py --deck /tmp/spsph_variant_deck --warmup 3 --steps 10 --profile > $out/profile_$name.log 2>&1
  head -1 $out/profile_$name.log | tee -a $out/summary.txt
done
cat $out/summary.txt
