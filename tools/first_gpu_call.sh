#!/bin/bash
# First GPU call of a round: everything that has to be confirmed on hardware before more is built on it.
#   gpurun --timeout 2400 -- 'bash tools/first_gpu_call.sh'
# 1. the device paths whose first hardware run is still pending (tests/test_zz_gpu_new_paths.py), one pytest
#    process per test id so that a crash in one cannot hide the others;
# 2. the verified GPU suite (-x as the round driver runs it);
# 3. the default bench line and the reference arm;
# 4. the ncu launch list of one step (shares per kernel);
# 5. the build-time variants of the pair-sum kernels (tools/variant_timing.sh; run `bash tools/variant_timing.sh build`
#    in the development container first so that the variant libraries travel with the snapshot).
# Everything lands in gpurun_out/first/.
out=gpurun_out/first
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
python -m pytest tests/test_zz_gpu_new_paths.py --collect-only -q 2>/dev/null | grep "::" > $out/new_ids.txt
while read -r id; do
  name=$(echo "$id" | sed 's/.*:://; s/[^A-Za-z0-9_]/_/g')
  timeout 900 python -m pytest "$id" -q -x > $out/new_$name.log 2>&1
  echo "$id: exit $?" | tee -a $out/new_paths_summary.txt
done < $out/new_ids.txt
timeout 3000 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_gpu_new_paths.py > $out/gpu_suite.log 2>&1
echo "gpu suite: exit $?" | tee -a $out/new_paths_summary.txt
tail -3 $out/gpu_suite.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
tail -c 1500 $out/bench.json
timeout 900 python bench.py --impl reference --steps 10 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 49 -c 60 --csv --log-file $out/launches.csv \
    python tools/run_steps.py --steps 2 > $out/launch_run.log 2>&1
bash tools/variant_timing.sh $out/variants > $out/variants.log 2>&1
cat $out/new_paths_summary.txt $out/variants/summary.txt
