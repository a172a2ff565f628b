"""Aggregate `ncu --page source --print-source cuda,sass --csv` by CUDA source line (top stall samples)."""
import csv, subprocess, sys, os
rep, kernel = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                     capture_output=True, text=True).stdout
fname = ''; hdr = None; agg = []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == 'File Path': fname = os.path.basename(r[1]); continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit():
        si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
        try: agg.append((int(r[si] or 0), int(r[ii] or 0), f"{fname}:{r[0]}", r[1].strip()[:100]))
        except ValueError: pass
tot = sum(a[0] for a in agg) or 1; toti = sum(a[1] for a in agg) or 1
print(f'kernel {kernel}: total samples {tot}, warp instructions {toti}')
for a in sorted(agg, reverse=True)[:topn]:
    print(f"{100*a[0]/tot:5.1f}% samp {100*a[1]/toti:5.1f}% inst | {a[2]:22s} {a[3]}")
