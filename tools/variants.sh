#!/bin/bash
# Build-time variants timed in ONE gpurun call (generalisation of tools/variant_timing.sh).
#   here (no GPU):  bash tools/variants.sh build "name:flags" "name2:flags2" ...   -> stress-particle-sph_b200/variants/libspsph_cuda_<name>.so
#   on the GPU box: bash tools/variants.sh run <outdir> [ENV=VAL ...] -- name name2 ...  -> parity (smoke) + 10-step event profile of the 4 M column
cd "$(dirname "$0")/.."
vdir=stress-particle-sph_b200/variants
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p $vdir
  for v in "$@"; do
    name=${v%%:*}; flags=${v#*:}
    ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -shared \
        -Xptxas -v $flags -Iinclude -Istress-particle-sph_b200/csrc -o $vdir/libspsph_cuda_$name.so \
        stress-particle-sph_b200/csrc/spsph_engine.cu -ldl > $vdir/ptxas_$name.log 2>&1 || echo "BUILD FAILED: $name" ) &
  done
  wait
  ls -la $vdir/*.so
  exit 0
fi
out=$1; shift
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
steps=${STEPS:-10}
for name in "$@"; do
  so=$PWD/$vdir/libspsph_cuda_$name.so
  [ -f $so ] || { echo "$name: library missing"; continue; }
  SPSPH_CUDA_SO=$so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/parity_$name.log 2>&1
  echo "$name parity exit $?" | tee -a $out/summary.txt
  SPSPH_CUDA_SO=$so timeout 600 python tools/run_steps.py --deck /tmp/spsph_variant_deck --warmup 3 --steps $steps --profile > $out/profile_$name.log 2>&1
  head -1 $out/profile_$name.log | cut -c1-80 | tee -a $out/summary.txt
  grep -E "k_count|k_fill|k_sweep_a|k_sweep_b|k_move|k_rank|k_scan" $out/profile_$name.log | awk '{printf "    %s %s", $1, $2} END {print ""}' | tee -a $out/summary.txt
done
