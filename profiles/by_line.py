"""Per CUDA source line: executed warp-instructions and stall samples of one kernel in an .ncu-rep
(needs -lineinfo and --import-source on).  usage: python profiles/by_line.py x.ncu-rep [min_pct]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(out.splitlines()))
lines, fname = [], ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].strip().isdigit() and r[7].strip().isdigit():
        lines.append((fname, int(r[0]), r[1].strip(), int(r[6]) if r[6].strip().isdigit() else 0, int(r[7])))     # samples, executed
te, ts = sum(l[4] for l in lines), sum(l[3] for l in lines)
print(f"total warp-instr {te}  samples {ts}")
for f, ln, src, sm, ex in lines:
    if ex > te * thr / 100 or sm > ts * thr / 100:
        print(f"{f[:14]:14s}{ln:5d} ex {100*ex/te:5.1f}% smp {100*sm/ts:5.1f}%  {src[:95]}")
