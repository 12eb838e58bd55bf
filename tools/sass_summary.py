"""Per-kernel SASS mnemonic counts of csrc/libgpb.so (cuobjdump -sass; runs without a GPU).
usage: python tools/sass_summary.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussianprocessfundamentals_b200", "csrc", "libgpb.so")
WANT = ["DMMA", "DFMA", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "RED", "LDL", "STL", "LDS", "LDG", "STG", "BAR", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", ln)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    counts = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for w in WANT:
                if op.startswith(w):
                    counts[cur][w] += 1
                    break
    print("# cuobjdump -sass / -res-usage of csrc/libgpb.so (sm_100a): instruction counts per kernel (static)")
    print("%-78s %5s %6s %5s %6s | %s" % ("kernel (demangled, shortened)", "regs", "smem", "local", "instr", " ".join("%6s" % w for w in WANT)))
    for name, c in counts.items():
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"gpb::", "", dem)
        dem = re.sub(r"GemmCfg<(\d+), (\d+), \d+, \d+, \d+>", r"Cfg\1x\2", dem)
        dem = re.sub(r"\(.*", "", dem).replace("void ", "")
        r = regs.get(name, (0, 0, 0))
        print("%-78s %5d %6d %5d %6d | %s" % (dem[:78], r[0], r[1], r[2], c["_total"], " ".join("%6d" % c[w] for w in WANT)))


if __name__ == "__main__":
    main()
