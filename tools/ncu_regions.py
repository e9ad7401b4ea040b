import csv, io, subprocess, sys, re
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],stdout=subprocess.PIPE,text=True).stdout
lines=raw.split("\n")
start=next(i for i,l in enumerate(lines) if l.startswith('"Address"'))
rows=list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base=int(rows[0]["Address"],16)
tot=sum(int(r["# Samples"] or 0) for r in rows)
exe=sum(int(r["Instructions Executed"] or 0) for r in rows)
addr=[int(r["Address"],16)-base for r in rows]
a2i={a:i for i,a in enumerate(addr)}
reasons=[k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
loops=[]
for i,r in enumerate(rows):
    src=r["Source"]
    m=re.search(r'\bBRA\b.*?(0x[0-9a-f]+)',src)
    if m:
        tgt=int(m.group(1),16)-base
        if tgt<=addr[i] and tgt in a2i:
            j=a2i[tgt]
            body=rows[j:i+1]
            nf=sum(1 for x in body if re.search(r'\bD(FMA|ADD|MUL)\b',x["Source"]))
            s=sum(int(x["# Samples"] or 0) for x in body)
            e=sum(int(x["Instructions Executed"] or 0) for x in body)
            loops.append((j,i,len(body),nf,s,e))
# innermost loops only: sort by length, mark covered
loops.sort(key=lambda x:x[2])
covered=[False]*len(rows)
print("total samples",tot,"executed",exe)
print("len nfp64  samples%  exec%   cycles/instr(rel)  range")
acc_s=acc_e=0
for j,i,n,nf,s,e in loops:
    if any(covered[j:i+1]): continue
    if s/tot<0.004: continue
    for k in range(j,i+1): covered[k]=True
    acc_s+=s; acc_e+=e
    why={}
    for k in reasons: why[k]=sum(int(x[k] or 0) for x in rows[j:i+1])
    top=sorted(why.items(),key=lambda kv:-kv[1])[:4]
    print(f"{n:4d} {nf:4d}   {100*s/tot:6.2f}   {100*e/exe:6.2f}   {(s/tot)/(e/exe):5.2f}   {hex(addr[j])}-{hex(addr[i])}  "+", ".join(f"{k[6:]}:{100*v/s:.0f}" for k,v in top))
print(f"innermost loops total: samples {100*acc_s/tot:.1f}%  executed {100*acc_e/exe:.1f}%")
rs=sum(int(r["# Samples"] or 0) for k,r in enumerate(rows) if not covered[k]); re_=sum(int(r["Instructions Executed"] or 0) for k,r in enumerate(rows) if not covered[k])
why={}
for k in reasons: why[k]=sum(int(r[k] or 0) for kk,r in enumerate(rows) if not covered[kk])
top=sorted(why.items(),key=lambda kv:-kv[1])[:6]
print(f"outside: samples {100*rs/tot:.1f}% executed {100*re_/exe:.1f}% ratio {(rs/tot)/(re_/exe):.2f} "+", ".join(f"{k[6:]}:{100*v/rs:.0f}" for k,v in top))
