"""Per-kernel SASS instruction census of libdqmc_b200.so (cuobjdump -sass):  python tools/sass_summary.py > profiles/sass_summary.txt
Columns: total instructions, DMMA (FP64 tensor core), DFMA/DADD/DMUL (vector FP64), LDGSTS (cp.async), UBLKCP/UTMALDG (TMA),
UTC*MMA (tcgen05: must be 0, it has no FP64 path), BAR (CTA barriers), SYNCS (mbarrier), LDS/STS, LDG/STG, spills (LDL/STL)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "dqmc_b200", "libdqmc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = [("DMMA", r"\bDMMA"), ("DFMA+", r"\bD(FMA|ADD|MUL)\b"), ("LDGSTS", r"\bLDGSTS"), ("TMA", r"\b(UBLKCP|UTMALDG|UTMASTG)"),
        ("UTCMMA", r"\bUTC\w*MMA"), ("BAR", r"\bBAR\."), ("SYNCS", r"\bSYNCS"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
        ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("spill", r"\b(LDL|STL)")]
kern, rows = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = dict(total=0, **{c: 0 for c, _ in cols})
        continue
    if kern and re.search(r"^\s+/\*[0-9a-f]{4,}\*/", line):
        rows[kern]["total"] += 1
        for c, pat in cols:
            if re.search(pat, line):
                rows[kern][c] += 1
arch = re.findall(r"arch = (\S+)", out)
print(f"# {os.path.relpath(so, ROOT)}: arch {sorted(set(arch))}, {len(rows)} kernels")
hdr = ["kernel", "total"] + [c for c, _ in cols]
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0][:70]
    except Exception:
        return n[:70]
for k, r in sorted(rows.items(), key=lambda kv: -kv[1]["total"]):
    print("| " + " | ".join([demangle(k), str(r["total"])] + [str(r[c]) for c, _ in cols]) + " |")
tot = {c: sum(r[c] for r in rows.values()) for c, _ in cols}
print("| **all kernels** | " + str(sum(r["total"] for r in rows.values())) + " | " + " | ".join(str(tot[c]) for c, _ in cols) + " |")
