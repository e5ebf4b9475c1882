"""Per-region view of an .ncu-rep source page: instructions in address order, aggregated in chunks, with executed counts,
stall samples and the dominant stall reasons -- shows WHERE in a long unrolled kernel the time goes."""
import csv, io, subprocess, sys
rep = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 200
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iSrc, iE = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[iS] or 0) for r in data) or 1
for a in range(0, len(data), chunk):
    seg = data[a:a + chunk]
    n = sum(int(r[iS] or 0) for r in seg)
    ex = max(int(r[iE] or 0) for r in seg)
    st = sorted(((sum(int(r[i] or 0) for r in seg), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    ops = {}
    for r in seg:
        w = r[iSrc].split()
        op = (w[1] if w and w[0].startswith('@') else (w[0] if w else '?')).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print(f"{a:6d}  samples {100*n/tot:5.1f}%  max exec {ex:10d}  {', '.join(f'{k}={100*v/max(n,1):.0f}%' for v, k in st):45s} {top}")
