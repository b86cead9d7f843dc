"""Aggregate ncu warp-stall samples per CUDA source line:  python tools/ncu_lines.py report.ncu-rep [top] [kernel-index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = {}; kern = None
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': kern = r[1][:40]; continue
    if len(r) > 5 and r[0].isdigit():
        try: agg[(kern, cur, int(r[0]))] = agg.get((kern, cur, int(r[0])), 0) + int(r[4]); src = r[1]
        except Exception: continue
        agg.setdefault(('src', kern, cur, int(r[0])), r[1][:100])
tot = {}
for k, v in agg.items():
    if k[0] != 'src': tot[k[0]] = tot.get(k[0], 0) + v
for kname, t in tot.items():
    print(f"== {kname}: {t} samples")
    items = sorted(((v, k) for k, v in agg.items() if k[0] == kname), reverse=True)[:top]
    for v, k in items:
        print(f"{v:7d} {v/t*100:5.1f}%  {k[1]}:{k[2]}  {agg[('src',)+k]}")
