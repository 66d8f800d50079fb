#!/bin/bash
# 2-GPU A/B of halo peeling on a 1 M-particle column (slab 170 cells, halo 13: the geometry of N = 4 on the 4 M column)
out=gpurun_out/peel
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
for v in 1 0; do
  extra=""; [ $v = 0 ] && extra="--no-extras"
  SPSPH_PEEL=$v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$v \
    bench.py --gpus 2 --no-cpu --ncol 816 --steps 30 --warmup 3 --weak-ncol 200 --weak-steps 4 --parity-ncol 136 --parity-steps 30 $extra \
    > $out/bench_peel$v.json 2> $out/bench_peel$v.err
  echo "peel=$v exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_peel$v.json").read().strip().split("\n")[-1])
    print("peel=$v ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "parity", (d.get("parity_check") or {}).get("bitwise_equal_single_gpu"))
except Exception as e:
    print("no line", e)
PY
  tail -2 $out/bench_peel$v.err
done
