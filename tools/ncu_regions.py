"""Developer tool: stall samples of an `ncu --page source --csv --print-source sass` export by SASS region (chunks of N
instructions) with the landmark instructions of each chunk. usage: ncu_regions.py file.csv [chunk]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if r and r[0].startswith("0x"):
        data.append(r)
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print('total samples', tot, 'exec', totex, 'n', len(data))
keys = ('UTCHMMA', 'LDTM', 'STTM', 'SYNCS', 'BAR.SYNC', 'UTCBAR', 'MEMBAR', 'FENCE', 'BAR.RED', 'CALL', 'RET', 'ELECT', 'MUFU.SQRT',
        'STG', 'LDG', 'MUFU.EX2', 'SHFL', 'DADD', 'LDS', 'STS', 'MUFU.RCP64H')
step = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for i in range(0, len(data), step):
    ch = data[i:i + step]
    s = sum(int(r[isamp]) for r in ch); e = sum(int(r[iex]) for r in ch)
    if s < tot * 0.003:
        continue
    st = {}
    for c in stall:
        v = sum(int(r[c]) for r in ch)
        if v:
            st[hdr[c][6:]] = v
    tst = [(k, round(100 * v / tot, 1)) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3]]
    lm = {}
    for r in ch:
        for k in keys:
            if k in r[isrc]:
                lm[k] = lm.get(k, 0) + 1
    print('%5d %5.1f%% exec %4.1f%%' % (i, 100 * s / tot, 100 * e / totex), tst, lm)
