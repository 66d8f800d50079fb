#!/bin/bash
# whole single-GPU suite + smoke + default bench line
out=${1:-gpurun_out/suite}
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > $out/gpu_suite.log 2>&1; echo "gpu suite exit $?"; tail -4 $out/gpu_suite.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/suite/bench.json").read().strip().split("\n")[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "roofline frac", d["roofline"]["frac"], "weak", d.get("weak_scaling",{}).get("ms_per_step"))
PY
