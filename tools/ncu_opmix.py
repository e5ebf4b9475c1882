import csv, subprocess, sys, io
rep=sys.argv[1]; launch=sys.argv[2] if len(sys.argv)>2 else "0"
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--launch-skip",launch,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=rows[1]; iSrc=hdr.index('Source'); iE=hdr.index('Instructions Executed')
opc={}; seen=set()
for r in rows[2:]:
    if r[0] in seen: continue
    seen.add(r[0])
    try: e=int(r[iE] or 0)
    except: continue
    t=r[iSrc].split()
    if not t: continue
    op=t[1] if t[0].startswith('@') else t[0]
    op='.'.join(op.split('.')[:2]) if op.startswith(('IMAD','LDS','LDG','STG','F2F','I2F','MUFU','SHFL')) else op.split('.')[0]
    opc[op]=opc.get(op,0)+e
ws=1.024e7/32
tot=sum(opc.values())
print("total warp-instr",tot,"per step",tot/ws)
for op,e in sorted(opc.items(),key=lambda kv:-kv[1]):
    if e/ws>0.2: print(f"{op:16s} {e/ws:7.2f}")
