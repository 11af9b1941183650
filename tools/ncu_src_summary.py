"""Developer tool: summarise an `ncu --page source --csv --print-source sass` export by code region."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print('total samples', tot, 'warp instrs executed', totex, 'static instrs', len(data))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 250
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(data)
def op(r):
    t = r[isrc].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
for i in range(lo, hi, step):
    chunk = data[i:i + step]
    s = sum(int(r[isamp]) for r in chunk); e = sum(int(r[iex]) for r in chunk)
    if s == 0 and e == 0: continue
    ops = {}
    for r in chunk: ops[op(r)] = ops.get(op(r), 0) + int(r[iex])
    st = {}
    for c in stall_cols:
        v = sum(int(r[c]) for r in chunk)
        if v: st[hdr[c][6:]] = v
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    tst = sorted(st.items(), key=lambda x: -x[1])[:4]
    print('%5d samples %5d (%4.1f%%) exec %9d (%4.1f%%) cyc/instr %5.1f' % (i, s, 100 * s / tot, e, 100 * e / totex, (s / tot) / max(e / totex, 1e-9)), top, tst)
