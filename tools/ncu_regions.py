import csv, subprocess, sys, io
rep=sys.argv[1]
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=rows[1]; iSrc=hdr.index('Source'); iS=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
data=[]; seen=set()
for r in rows[2:]:
    if r[0] in seen: continue
    seen.add(r[0])
    try: data.append((r[iSrc], int(r[iS] or 0), int(r[iE] or 0)))
    except: pass
tot=sum(d[1] for d in data); tote=sum(d[2] for d in data)
print("instructions",len(data),"samples",tot,"executed",tote)
# markers
marks=[i for i,d in enumerate(data) if any(k in d[0] for k in ("LDGDEPBAR","DEPBAR","BAR.SYNC","SYNCS","ATOM","RED","EXIT","WARPSYNC"))]
prev=0
for m in marks+[len(data)]:
    s=sum(d[1] for d in data[prev:m]); e=sum(d[2] for d in data[prev:m])
    if m>prev: print(f"[{prev:5d},{m:5d})  samples {100*s/tot:5.1f}%  exec {100*e/tote:5.1f}%   next: {data[m][0][:60] if m<len(data) else 'END'}")
    prev=m
