"""Count the Blackwell-native SASS opcodes per kernel of the built libnerftiny.so (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP; HMMA would
be the legacy mma.sync path.

    python tools/sass_opcodes.py > profiles/sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nerf_tiny_b200", "libnerftiny.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.split("\n")
    counts, order, cur, k = collections.OrderedDict(), [], None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[k].replace("(anonymous namespace)::", "")).replace("void ", "")
            k += 1
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        counts[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
        counts[cur]["_total"] += 1
    print("# SASS opcode counts per kernel of nerf_tiny_b200/libnerftiny.so (cuobjdump -sass, sm_100a); tools/sass_opcodes.py")
    print("# UTCHMMA = tcgen05.mma.kind::f16 (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor")
    print("# load/store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (none)")
    print("%-64s %s %8s" % ("kernel", " ".join("%12s" % o for o in OPS), "total"))
    tot = collections.Counter()
    for name, c in counts.items():
        print("%-64s %s %8d" % (name[:64], " ".join("%12d" % c.get(o, 0) for o in OPS), c["_total"]))
        tot.update(c)
    print("%-64s %s %8d" % ("ALL KERNELS", " ".join("%12d" % tot.get(o, 0) for o in OPS), tot["_total"]))


if __name__ == "__main__":
    sys.exit(main())
