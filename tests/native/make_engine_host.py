#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Rewrites csrc/spsph_engine.cu into a translation unit g++ can compile for the serial host
emulation (tests/native/cuda_host_emu.h): the only non-C++ construct in the file is the kernel launch syntax,
    kernel<<<grid, block, shared, stream>>>(args...)   ->   emu_launch(grid, block, kernel, args...)
usage: make_engine_host.py <spsph_engine.cu> <out.cpp>"""
import re
import sys

src = open(sys.argv[1]).read()
pat = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<\s*(.*?)\s*>>>\s*\(", re.S)


def repl(m):
    cfg = [c.strip() for c in m.group(2).split(",")]
    # grid and block may themselves contain commas only inside parentheses: re-join by depth
    parts, depth, cur = [], 0, ""
    for ch in m.group(2):
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
            continue
        depth += ch in "([{"
        depth -= ch in ")]}"
        cur += ch
    parts.append(cur.strip())
    return f"emu_launch({parts[0]}, {parts[1]}, {m.group(1)}, "


out, n = pat.subn(repl, src)
open(sys.argv[2], "w").write('#define SPSPH_EMU_RUNTIME 1\n#include "cuda_host_emu.h"\n' + out)
print(f"{n} kernel launches rewritten")
