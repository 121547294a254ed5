"""Top stalled SASS instructions of a kernel in an .ncu-rep (source page)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ia, isamp = hdr.index('Source'), hdr.index('# Samples')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print('kernel:', rows[0][1][:100] if rows[0] else '', ' total samples', tot, ' sass lines', len(data))
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:topn]:
    st = sorted([(int(r[i] or 0), h) for i, h in stall], reverse=True)[:2]
    print(f"{int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  {r[ia][:64]:64s} {st}")
agg = {}
for r in data:
    for i, h in stall:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print({k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]})
