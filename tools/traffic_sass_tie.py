#!/usr/bin/env python
"""Ties the ncu traffic capture under profiles/sweep_b_traffic.json to the MACHINE CODE of the kernels it measured.

The capture records a hash of the csrc/ sources it was taken from; any later edit of those files -- a comment, host
code, a new unrelated kernel -- would detach it although the measured kernels are unchanged. This tool rebuilds the
library from the sources of the capture's commit (git show <commit>:csrc/*, checked against the recorded source hash),
hashes the SASS of every instantiation of the captured kernels and stores the hashes in the json. bench.py attaches
`roofline.traffic` when the library it timed carries the same SASS for all of them (or the source hash still matches).

usage: python tools/traffic_sass_tie.py <commit of the capture>      (here, no GPU needed)"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as g  # noqa: E402
from sass_guard_check import sass_by_function  # noqa: E402

CSRC = "stress-particle-sph_b200/csrc"


def main(commit):
    path = os.path.join(ROOT, "profiles", "sweep_b_traffic.json")
    tj = json.load(open(path))
    tmp = tempfile.mkdtemp(prefix="spsph_tie_")
    files = subprocess.run(["git", "ls-tree", "--name-only", commit, CSRC + "/"], cwd=ROOT, check=True,
                           stdout=subprocess.PIPE, text=True).stdout.split()
    h = hashlib.sha256()
    for f in sorted(files, key=os.path.basename):
        blob = subprocess.run(["git", "show", f"{commit}:{f}"], cwd=ROOT, check=True, stdout=subprocess.PIPE).stdout
        open(os.path.join(tmp, os.path.basename(f)), "wb").write(blob)
        h.update(os.path.basename(f).encode())
        h.update(blob)
    assert h.hexdigest() == tj["csrc_sha256"], f"{commit} is not the commit of the capture: csrc hash {h.hexdigest()[:12]}"
    inc = os.path.join(tmp, "include")
    os.makedirs(inc)
    open(os.path.join(inc, "spsph.h"), "wb").write(
        subprocess.run(["git", "show", f"{commit}:include/spsph.h"], cwd=ROOT, check=True, stdout=subprocess.PIPE).stdout)
    so = os.path.join(tmp, "libspsph_cuda_capture.so")
    g._run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + g.NVCC_FLAGS +
           ["-I" + inc, "-I" + tmp, "-o", so, os.path.join(tmp, "spsph_engine.cu"), "-ldl"])
    sass = sass_by_function(so)
    names = sorted(tj["kernels"])
    tie = {m: v for m, v in sass.items() if any(f"{len(n)}{n}" in m for n in names)}  # mangled: <len><name>
    assert tie and all(any(f"{len(n)}{n}" in m for m in tie) for n in names), "a captured kernel is missing"
    tj["kernel_sass_sha256"] = tie
    tj["sass_tie"] = (f"SASS hashes (cuobjdump -sass, per mangled function, address and encoding columns included) of every "
                      f"instantiation of the captured kernels, from a rebuild of commit {commit} (= csrc_sha256) with "
                      f"build()'s flags; tools/traffic_sass_tie.py")
    json.dump(tj, open(path, "w"), indent=1)
    print(f"{len(tie)} kernel instantiations tied:", ", ".join(names))


if __name__ == "__main__":
    main(sys.argv[1])
