#!/bin/bash
# Memory checks of the device code WITHOUT a GPU, on the serial host emulation of the engine
# (tests/native/cuda_host_emu.h): the stand-ins for compute-sanitizer's memcheck and initcheck.
#   1. AddressSanitizer build of the emulated engine ("device" allocations are exact-size heap blocks, so an
#      out-of-bounds list / state access of any per-particle kernel is a heap-buffer-overflow report);
#   2. SPSPH_EMU_POISON=1: fresh "device" memory holds 0xFF bytes instead of zeros; the reference-parity suite must
#      still pass bit for bit, i.e. no kernel depends on memory nobody wrote.
# usage: bash tools/emulated_sanitizers.sh   (about 4 minutes)
#        EMU_FLAGS="-DSPSPH_ELL_PIPE=1 -DSPSPH_ELL_SUB=2 -DSPSPH_A_SUB=2" SIMT_ONLY=1 bash tools/emulated_sanitizers.sh
#        (a build-time variant of tools/variant_timing.sh under the SIMT emulation only)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
python "$root/tests/native/make_engine_host.py" "$root/stress-particle-sph_b200/csrc/spsph_engine.cu" "$tmp/engine_host.cpp"
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -fno-fast-math -std=c++17 -fPIC -shared -w \
    -D__noinline__= $EMU_FLAGS -I/usr/local/cuda/include -I"$root/tests/native" -I"$root/stress-particle-sph_b200/csrc" \
    -I"$root/include" -o "$tmp/libspsph_asan.so" "$tmp/engine_host.cpp" -ldl
cat > "$tmp/run.py" <<PY
import sys, tempfile
sys.path.insert(0, "$root/stress-particle-sph_b200"); sys.path.insert(0, "$root/oracle")
import spsph.engine as E
E._lib, E._CUDA_SO = None, "$tmp/libspsph_asan.so"
import spsph
from spsph import decks
from ref_cases import spec_of
for case, n in (("bui", 30), ("vs", 20), ("sl", 5), ("bui_inside_sp1", 20), ("bui_standard", 10), ("sl_sigman_xsph", 10),
                ("bui_cont_density", 8), ("bui_art_stress", 6), ("bui_sml15", 6), ("bui_out_domain", 30), ("bui_long", 310)):
    variant, spec = spec_of(case)
    d = tempfile.mkdtemp(); decks.write_deck(d, spec); prob = spsph.load(d, variant)
    eng = spsph.Engine(prob); eng.run(1, 0.0, prob.blocks[0]["dt"], n); eng.download(); eng.pairs(); eng.pair_stats(); eng.close()
    print("asan:", case, n, "steps clean", flush=True)
PY
[ -n "$SIMT_ONLY" ] || LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python "$tmp/run.py"
# 1b. the same under the lockstep (SIMT) emulation: the real cp.async list streaming (which reads whole groups of rows,
#     i.e. past the end of a slice into the allocation slack), shuffles and block scans
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -fno-fast-math -std=c++17 -fPIC -shared -w \
    -DSPSPH_EMU_SIMT -D__noinline__= $EMU_FLAGS -I/usr/local/cuda/include -I"$root/tests/native" \
    -I"$root/stress-particle-sph_b200/csrc" -I"$root/include" -o "$tmp/libspsph_asan.so" "$tmp/engine_host.cpp" -ldl
sed -i 's/("bui", 30), ("vs", 20), ("sl", 5), ("bui_inside_sp1", 20), ("bui_standard", 10), ("sl_sigman_xsph", 10),/("bui", 4), ("vs", 4), ("bui_inside_sp1", 4), ("sl_sigman_xsph", 3),/; s/("bui_cont_density", 8), ("bui_art_stress", 6), ("bui_sml15", 6), ("bui_out_domain", 30), ("bui_long", 310)):/("bui_sml15", 3), ("bui_out_domain", 4)):/' "$tmp/run.py"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python "$tmp/run.py"
[ -n "$SIMT_ONLY" ] || SPSPH_EMU_POISON=1 python -m pytest "$root/tests/test_engine_emulated_reference_cpu.py" -q -x
