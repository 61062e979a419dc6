"""Blackwell opcode evidence: per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics in the built library
(`cuobjdump -sass`), so that the proof is tracked in the repository and not only inside an untracked .so or .ncu-rep.

    python profiles/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "birdsoundclassif_b200", "libnbm_b200.so")
# UTCHMMA = tcgen05.mma (f16 kind), UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UTCATOMSWS = tcgen05.alloc,
# UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch, SYNCS = mbarrier ops, ELECT = elect.sync, MUFU.LG2 = lg2.approx
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UBLKPF", "SYNCS", "ELECT", "MUFU.LG2", "FMNMX3",
         "DFMA", "ATOM", "ATOMG", "RED", "SHFL", "BAR.SYNC", "HADD2", "PRMT", "LDS", "STS", "LDG", "STG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel")
    print(f"{'kernel':46s} {'total':>7s} " + " ".join(f"{w:>9s}" for w in WATCH))
    for k, c in kernels.items():
        print(f"{k[-46:]:46s} {c['_total']:7d} " + " ".join(f"{c[w]:9d}" for w in WATCH))


if __name__ == "__main__":
    sys.exit(main())
