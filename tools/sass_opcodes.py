"""Per-kernel count of the SASS opcodes that prove what hardware a kernel uses (tcgen05 MMA / TMEM / TMA / mbarrier), from
`cuobjdump -sass` of the built library.  Runs where the CUDA toolkit is installed (no GPU needed).
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "knowledge-distillation-by-replacing-cheap-conv_b200", "libkdcc.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "ELECT",
       "FFMA", "FFMA2", "HFMA2", "SHFL", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("kdcc::", "")
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            base = op.split(".")[0]
            counts[cur][base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                counts[cur]["UTCHMMA.2CTA"] += 1
            counts[cur]["_total"] += 1
    head = git = ""
    try:
        head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except Exception:
        pass
    print("# SASS opcode counts per kernel of libkdcc.so (cuobjdump -sass, sm_100a), summarised at commit %s" % head)
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st (TMEM), UTCCP = tcgen05.cp, UTCBAR = tcgen05.commit,")
    print("# UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier ops, ELECT = elect.sync")
    used = [o for o in OPS if any(counts[k][o] for k in order)]
    print("%-64s %7s " % ("kernel", "instrs") + " ".join("%8s" % o[:8] for o in used))
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        print("%-64s %7d " % (k[:64], c["_total"]) + " ".join("%8d" % c[o] for o in used))
        tot.update(c)
    print("%-64s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[o] for o in used))


if __name__ == "__main__":
    sys.exit(main())
