"""Developer tool: per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses
(cuobjdump -sass of the shipped libtsp_b200.so): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor
load (cp.async.bulk.tensor), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, HMMA = legacy mma.sync, REDUX."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tree_search_planning_b200", "libtsp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, universal_newlines=True, check=True).stdout
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "REDUX", "FENCE.VIEW.ASYNC", "ATOMS", "DADD", "DMUL"]
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, universal_newlines=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k in MN:
            if re.search(r"\b" + re.escape(k), line):
                counts[cur][k] += 1
print("%-62s " % "kernel (libtsp_b200.so, sm_100a)" + " ".join("%8s" % k[:8] for k in MN))
for name, c in counts.items():
    if name.startswith("tsp::") or "kernel" in name:
        print("%-62s " % name[:62] + " ".join("%8d" % c[k] for k in MN))
