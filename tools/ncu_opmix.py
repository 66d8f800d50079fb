#!/usr/bin/env python
"""instruction mix of an .ncu-rep (--set full --import-source on): warp instructions per opcode class and per
execution-count bucket (executions per warp), to see where a kernel's issue slots go"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, warps=None):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    idx = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[2:] if len(r) > idx["# Samples"] and r[idx["# Samples"]].isdigit()]
    ex = [int(r[idx["Instructions Executed"]]) for r in data]
    tot = sum(ex)
    W = warps or ex[0]  # the first instruction runs once per warp
    print(rows[0][1][:80])
    print(f"warps {W}, warp instructions {tot/1e6:.1f} M = {tot/W:.0f} per warp")
    ops = collections.Counter()
    cls = collections.Counter()
    for r, n in zip(data, ex):
        t = r[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += n
        c = ("fp64" if op in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU") else
             "fp32" if op in ("FFMA", "FMUL", "FADD", "FSETP", "FMNMX", "FSEL") else
             "cvt" if op in ("F2F", "I2F", "F2I", "I2FP") else
             "ldst" if op in ("LDG", "STG", "LDS", "STS", "LD", "ST", "LDGSTS", "LDC", "LDCU", "LDL", "STL", "LDGDEPBAR", "DEPBAR") else
             "ctrl" if op in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BAR", "WARPSYNC", "BREAK", "NOP", "YIELD") else
             "int/move")
        cls[c] += n
    print("  classes:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in cls.most_common()))
    print("  top ops:", ", ".join(f"{k} {v/W:.0f}" for k, v in ops.most_common(22)), "(per warp)")
    b = collections.Counter()
    for n in ex:
        b[round(n / W, 1)] += n
    print("  by executions per warp:", ", ".join(f"x{k:g}: {v/W:.0f}" for k, v in sorted(b.items(), key=lambda x: -x[1])[:10]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
