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
  "base:"                                               # every list staged in shared memory (LG_SP0=7 LG_N0=10 LG_NC=7 LG_SS=11)
  "n0reg:-DSPSPH_LG_N0=0"                               # velocity-particle side of sweeps A / B: lists read two groups ahead into registers
  "n0ncreg:-DSPSPH_LG_N0=0 -DSPSPH_LG_NC=0"             # ... and the artificial-viscosity list
  "allreg:-DSPSPH_LG_N0=0 -DSPSPH_LG_NC=0 -DSPSPH_LG_SP0=0 -DSPSPH_LG_SS=0"  # nothing staged but the partner tiles
  "w20:-DSPSPH_TILE_WARPS=20"                           # 20 resident warps per SM requested (102 registers)
  "w24:-DSPSPH_TILE_WARPS=24 -DSPSPH_LG_N0=0 -DSPSPH_LG_NC=0"  # 24 warps (80 registers)
  "list:"                                               # same library as base, run with SPSPH_TILE=0: the id-list path
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
  [ $name = list ] && export SPSPH_TILE=0
  SPSPH_CUDA_SO=$so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/parity_$name.log 2>&1
  echo "$name parity exit $?" | tee -a $out/summary.txt
  SPSPH_CUDA_SO=$so timeout 600 python tools/run_steps.py --deck /tmp/spsph_variant_deck --warmup 3 --steps 10 --profile > $out/profile_$name.log 2>&1
  head -1 $out/profile_$name.log | tee -a $out/summary.txt
  grep -E "k_sweep_a|k_sweep_b|k_move|k_tile_build" $out/profile_$name.log | awk '{printf "    %s %s", $1, $2} END {print ""}' | tee -a $out/summary.txt
  unset SPSPH_TILE
done
cat $out/summary.txt
