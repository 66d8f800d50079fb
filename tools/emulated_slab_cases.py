#!/usr/bin/env python
"""Any case of oracle/ref_cases.py on N host-emulated slab ranks (no GPU): builds the emulated engine and the NCCL
stand-in like tests/test_dist_emulated_cpu.py, runs tools/dist_worker.py under torch.distributed.run and compares the
merged owned particles and the all-reduced pair count with the single-domain oracle, bit for bit.

usage: python tools/emulated_slab_cases.py case:steps:ranks [case:steps:ranks ...]
  e.g. python tools/emulated_slab_cases.py bui_standard:30:2 bui_inside_sp1_long:1501:2 refined_bui@408:6:8
(round 2: all of bui_standard, bui_inside_sp1/3, bui_shift5, bui_quintic, bui_art_stress, sl_tresca, vs_standard,
bui_plane_stress, bui_sml15, bui_out_domain pass on two ranks; refined_bui@204:1200:4 -- 21 list-growth steps on four
ranks -- and refined_bui@408:6:8 pass as well)"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
sys.path[:0] = [os.path.join(ROOT, "stress-particle-sph_b200"), os.path.join(ROOT, "oracle")]
KEYS = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ")


def build(d):
    cpp, so, nccl = os.path.join(d, "engine_host.cpp"), os.path.join(d, "libspsph_emu.so"), os.path.join(d, "libfake_nccl.so")
    subprocess.run([sys.executable, os.path.join(NATIVE, "make_engine_host.py"),
                    os.path.join(ROOT, "stress-particle-sph_b200", "csrc", "spsph_engine.cu"), cpp], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                    "-D__noinline__=", "-fno-gnu-unique", "-I/usr/local/cuda/include", "-I" + NATIVE,
                    "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                    "-o", so, cpp, "-ldl"], check=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", nccl,
                    os.path.join(NATIVE, "fake_nccl.cpp")], check=True)
    return so, nccl


def main(argv):
    import spsph
    from spsph import decks, dist
    from ref_cases import spec_of
    from oracle_binding import Oracle
    work = tempfile.mkdtemp(prefix="spsph_emu_slabs_")
    so, nccl = build(work)
    env = dict(os.environ, SPSPH_EMU_SO=so, SPSPH_NCCL_SO=nccl, SPSPH_FAKE_NCCL_DIR=work, OMP_NUM_THREADS="1")
    port, failed = 29800, 0
    for arg in argv:
        case, steps, world = arg.split(":")
        steps, world, port = int(steps), int(world), port + 1
        out = tempfile.mkdtemp(dir=work)
        if case.startswith("refined_bui@"):  # the bench workload at a reduced size: refined_bui@<columns>
            ncol = int(case.split("@")[1])
            kind_args = ["--kind", "refined_bui", "--ncol", str(ncol)]
        else:
            kind_args = ["--kind", "case:" + case]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", str(port),
                            os.path.join(ROOT, "tools", "dist_worker.py")] + kind_args + ["--steps", str(steps),
                            "--out", out], capture_output=True, text=True, env=env)
        if r.returncode:
            print(case, "RUN FAILED:", [ln for ln in r.stderr.splitlines() if "Error" in ln][:3], flush=True)
            failed += 1
            continue
        variant, spec = ("bui", decks.refined_bui_spec(ncol=ncol)) if case.startswith("refined_bui@") else spec_of(case)
        d = tempfile.mkdtemp(dir=work)
        decks.write_deck(d, spec)
        prob = spsph.load(d, variant)
        orc = Oracle(prob)
        orc.run(1, 0.0, prob.blocks[0]["dt"], steps)
        ref = orc.download()
        ranks = [np.load(os.path.join(out, f"rank{k}.npz")) for k in range(world)]
        merged = dist.merge_owned([{k: r_[k] for k in KEYS} for r_ in ranks], [r_["flags"] for r_ in ranks], prob.params)
        nt = prob.params.ntotal
        bad = [k for k in KEYS if not np.array_equal(merged[k][:nt] if k in ("x", "vel", "stress") else merged[k],
                                                      ref[k][:nt] if k in ("x", "vel", "stress") else ref[k])]
        if int(ranks[0]["npairs"]) != orc.pair_stats()["npairs"]:
            bad.append("npairs")
        print(f"{case}: {steps} steps on {world} emulated ranks:", "bitwise equal to the oracle" if not bad else f"DIFFERS in {bad}",
              flush=True)
        failed += bool(bad)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
