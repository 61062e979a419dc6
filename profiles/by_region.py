"""Per pipeline phase of slide_ws_kernel: executed instructions, stall samples and top stall reasons.
usage: python profiles/by_region.py x.ncu-rep   (needs -lineinfo, --import-source on)"""
import csv, re, subprocess, sys, os
from collections import defaultdict, Counter
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
names = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_not_selected", "stall_math", "stall_mio", "stall_lg",
         "stall_branch_resolving", "stall_no_inst", "stall_dispatch", "stall_sleep", "stall_selected"]
idx = {n: hdr.index(n) for n in names}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "birdsoundclassif_b200/csrc/frontend_tc.cu")).read().splitlines()
pat = r"MMA issuer \(all|fill warp of group|worker group g ====|auto planes|auto build_unit|auto build_tail|auto build = |auto anchor_of|// ---- recur|// ---- next chain|// ---- emit|stage free for the next|float2 anc_next"
marks = [(i + 1, l.strip()[:34]) for i, l in enumerate(src) if re.search(pat, l)]
k0 = next(i + 1 for i, l in enumerate(src) if "slide_ws_body(const TcParams" in l)
def region(ln):
    name = "setup"
    for m, t in marks:
        if ln >= m: name = t
    return name
reg, ex = defaultdict(Counter), Counter()
cur, fname = None, ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0].strip().isdigit(): cur = (fname, int(r[0])); continue
    if len(r) > 8 and r[2].startswith("0x") and r[6].isdigit():
        name = region(cur[1]) if cur[0] == "frontend_tc.cu" and cur[1] > k0 else ("helpers/" + (("L%d" % cur[1]) if cur[0] == "frontend_tc.cu" else cur[0][:10]))
        for n, i in idx.items():
            if r[i].isdigit(): reg[name][n] += int(r[i])
        ex[name] += int(r[7])
tot = sum(sum(c.values()) for c in reg.values()); te = sum(ex.values())
print(f"{'region':36s} {'ex%':>5s} {'smp%':>5s}  top stalls")
for name, c in sorted(reg.items(), key=lambda x: -ex[x[0]]):
    sm = sum(c.values())
    if ex[name] < te * 0.003 and sm < tot * 0.003: continue
    print(f"{name:36s} {100*ex[name]/te:5.1f} {100*sm/max(tot,1):5.1f}  " + ", ".join(f"{k[6:]}:{100*v/max(sm,1):.0f}%" for k, v in c.most_common(5)))
