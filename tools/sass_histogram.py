"""Developer tool (CPU): opcode histogram of libxmm_b200.so per kernel -- the SASS mnemonics that prove the Blackwell
path (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UTMAPF = TMA
L2 prefetch, UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier).  python tools/sass_histogram.py > profiles/...txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "xmm_superres_denoise_b200", "libxmm_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
    print("%-78s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
    tot = collections.Counter()
    for name, c in per.items():
        d = re.sub(r"\(.*", "", demangle(name)).replace("void ", "").replace("xmm::", "")
        print("%-78s %7d " % (d[:78], c["total"]) + " ".join("%7d" % c[k] for k in KEYS))
        tot.update(c)
    print("%-78s %7d " % ("TOTAL (%d kernels)" % len(per), tot["total"]) + " ".join("%7d" % tot[k] for k in KEYS))


if __name__ == "__main__":
    main()
