#!/usr/bin/env python
"""Proof that the host-emulation alternatives inside csrc/ (#ifdef SPSPH_HOST_EMU / SPSPH_EMU_*: test scaffolding of
tests/native/) do not reach the product: the kernel sources are copied with every such conditional resolved as nvcc
sees it (macros undefined: emulation branches deleted, device branches kept without their guards), both trees are
compiled with build()'s flags, and the SASS of the two libraries is compared function by function.

usage: python tools/sass_guard_check.py [--out profiles/r2_sass_guard_check.txt]
also used by tests/test_oracle_cpu.py::test_sass_identical_without_emulation_guards (SPSPH_SASS_CHECK=1)."""
import hashlib
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "stress-particle-sph_b200", "csrc")
EMU = re.compile(r"SPSPH_HOST_EMU|SPSPH_EMU_\w+")


def _eval_guard(directive, expr):
    """value of an emulation guard when none of the emulation macros is defined"""
    if directive == "ifdef":
        return False
    if directive == "ifndef":
        return True
    e = re.sub(r"defined\s*\(\s*(SPSPH_HOST_EMU|SPSPH_EMU_\w+)\s*\)", "0", expr)
    e = re.sub(r"defined\s+(SPSPH_HOST_EMU|SPSPH_EMU_\w+)", "0", e)
    if EMU.search(e) or not re.fullmatch(r"[01\s!&|()]*", e):
        raise ValueError(f"guard too complex to resolve: #if {expr}")
    return bool(eval(e.replace("&&", " and ").replace("||", " or ").replace("!", " not ")))


def strip_guards(text):
    """-> (text without the emulation conditionals, number of guards resolved)"""
    out, stack, guards = [], [], 0  # stack of (is_emu_guard, branch_taken_now, parent_active)
    active = True
    for line in text.split("\n"):
        m = re.match(r"\s*#\s*(ifdef|ifndef|if|elif|else|endif)\b(.*)", line)
        if not m:
            if active:
                out.append(line)
            continue
        d, rest = m.group(1), m.group(2).split("//")[0].strip()
        if d in ("ifdef", "ifndef", "if"):
            emu = bool(EMU.search(rest))
            if emu:
                guards += 1
                val = _eval_guard(d, rest)
                stack.append((True, val, active))
                active = active and val
            else:
                stack.append((False, True, active))
                if active:
                    out.append(line)
        elif d == "elif":
            if stack[-1][0]:
                raise ValueError("#elif on an emulation guard is not supported")
            if active:
                out.append(line)
        elif d == "else":
            emu, val, parent = stack[-1]
            if emu:
                stack[-1] = (True, not val, parent)
                active = parent and not val
            elif active:
                out.append(line)
        else:  # endif
            emu, val, parent = stack.pop()
            if not emu and active:
                out.append(line)
            active = parent
    assert not stack, "unbalanced conditionals"
    return "\n".join(out), guards


def sass_by_function(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], check=True, stdout=subprocess.PIPE, text=True).stdout
    funcs, name, body = {}, None, []
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = hashlib.sha256("\n".join(body).encode()).hexdigest()
            # anonymous-namespace symbols carry a hash of the source path: not part of the comparison
            name, body = re.sub(r"_GLOBAL__N__[0-9a-f]{8}_", "_GLOBAL__N__", m.group(1)), []
        elif name and "/*" in line:
            body.append(re.sub(r"\s+", " ", line.strip()))
    if name:
        funcs[name] = hashlib.sha256("\n".join(body).encode()).hexdigest()
    return funcs


def check(report=None):
    import __graft_entry__ as g
    g.build()
    product = os.path.join(g.PKG, "libspsph_cuda.so")
    tmp = tempfile.mkdtemp(prefix="spsph_sass_guard_")
    try:
        n_guards = 0
        for f in sorted(os.listdir(CSRC)):
            stripped, n = strip_guards(open(os.path.join(CSRC, f)).read())
            if EMU.search(re.sub(r"//.*", "", stripped)):
                raise AssertionError(f"{f}: an emulation macro survives outside comments")
            n_guards += n
            open(os.path.join(tmp, f), "w").write(stripped)
        so = os.path.join(tmp, "libspsph_cuda_noguards.so")
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        flags = [x for x in g.NVCC_FLAGS if x != "-lineinfo"] + ["-lineinfo"]
        g._run([nvcc] + flags + ["-I" + os.path.join(ROOT, "include"), "-I" + tmp, "-o", so,
                                 os.path.join(tmp, "spsph_engine.cu"), "-ldl"])
        a, b = sass_by_function(product), sass_by_function(so)
        differ = sorted(k for k in set(a) | set(b) if a.get(k) != b.get(k))
        whole = hashlib.sha256("".join(f"{k}:{a[k]}\n" for k in sorted(a)).encode()).hexdigest()
        lines = [
            "SASS of libspsph_cuda.so built from csrc/ as committed vs. built from csrc/ with every",
            "#if(n)def SPSPH_HOST_EMU / SPSPH_EMU_* conditional resolved and removed (tools/sass_guard_check.py)",
            f"emulation guards resolved: {n_guards}",
            f"device functions compared: {len(a)} (product) / {len(b)} (without guards)",
            f"functions whose SASS differs: {len(differ)}" + ("".join("\n  " + d for d in differ[:20])),
            f"sha256 over the per-function SASS hashes of the product: {whole}",
            "RESULT: " + ("IDENTICAL" if not differ and a else "DIFFERENT"),
        ]
        if report:
            open(report, "w").write("\n".join(lines) + "\n")
        return (not differ and bool(a)), "\n".join(lines)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    ok, text = check(out)
    print(text)
    sys.exit(0 if ok else 1)
