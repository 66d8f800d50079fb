#!/bin/bash
# last call of the round: smoke + the default bench line from the final build
out=gpurun_out/final
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; tail -2 $out/bench.err
tail -c 3000 $out/bench.json
