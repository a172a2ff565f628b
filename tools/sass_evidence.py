"""Static evidence from the built library (no GPU needed): per kernel the SASS instruction count and the mnemonics that
matter for the parity / bandwidth claims — no fused multiply-add in the kernels whose bits are compared with
numpy / scipy, vector (128 / 256-bit, non-coherent) loads in the streaming kernels, no local-memory spills.
    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ortho2align_b200", "lib", "libo2a_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, stats = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern = kern.replace("o2a::", "").replace("void ", "")
        stats[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        c = stats[kern]
        c["instructions"] += 1
        base = op.split(".")[0]
        if base in ("DFMA", "FFMA"):
            c[base] += 1
        if base in ("DADD", "DMUL", "DSETP", "MUFU"):
            c[base] += 1
        if base == "LDG":
            width = "256" if ".256" in op else "128" if ".128" in op else "64" if ".64" in op else "32/less"
            c["LDG." + width + (".CONSTANT" if "CONSTANT" in op else "")] += 1
        if base in ("LDL", "STL"):
            c[base] += 1
        if base in ("ATOMS", "RED", "ATOMG", "SHFL", "VOTE", "MATCH"):
            c[base] += 1
print("SASS evidence of ortho2align_b200/lib/libo2a_cuda.so (sm_100a; nvcc -fmad=false; tools/sass_evidence.py)")
print("parity-critical kernels must show DFMA = 0 except inside erfc / division sequences of the CUDA math library")
print("(KDE branch, k_fit / k_pq / k_front<*, 1>: tolerance contract) — the histogram path compares bits.\n")
cols = ["instructions", "DFMA", "FFMA", "DADD", "DMUL", "MUFU", "LDG.256.CONSTANT", "LDG.128.CONSTANT", "LDG.128", "LDG.64", "LDG.64.CONSTANT",
        "LDG.32/less", "LDL", "STL", "ATOMS", "SHFL", "VOTE", "MATCH"]
print(f"{'kernel':58s} " + " ".join(f"{c[:11]:>11s}" for c in cols))
for k, c in stats.items():
    print(f"{k[:58]:58s} " + " ".join(f"{c.get(col, 0):11d}" for col in cols))
