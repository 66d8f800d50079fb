#!/bin/bash
out=gpurun_out/r2h
mkdir -p $out
python __graft_entry__.py > $out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tile_path.py -m gpu -q -k "restart or row_wise or long_run" > $out/tile_tests.log 2>&1
echo "tile tests: exit $?"; tail -3 $out/tile_tests.log
timeout 600 python tools/strict_cases.py sl_sigman vs_sigman sl_sigman_xsph sl_xsph > $out/strict.log 2>&1; cat $out/strict.log | tail -6
timeout 900 python -m pytest tests/test_full_size.py -m gpu -q -k 1m_20 > $out/one_m.log 2>&1; echo "1M test: exit $?"; tail -2 $out/one_m.log
for mode in 0 1; do
  for tool in memcheck racecheck initcheck; do
    SPSPH_TILE=$mode timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/run_steps.py --kind bui --steps 3 \
        > $out/sanitizer_${tool}_tile$mode.log 2>&1
    echo "sanitizer $tool tile=$mode: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tile/list" $out/sanitizer_${tool}_tile$mode.log | tail -3
  done
done
