import re,sys
lines=open(sys.argv[1]).read().splitlines()
ins=[]
for l in lines:
    m=re.search(r'/\*([0-9a-f]+)\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
addr2idx={a:i for i,(a,_) in enumerate(ins)}
loops=[]
for i,(a,t) in enumerate(ins):
    m=re.search(r'\bBRA\b.*?(0x[0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<=a and tgt in addr2idx:
            j=addr2idx[tgt]
            body=ins[j:i+1]
            nf=sum(1 for _,x in body if re.search(r'\bD(FMA|ADD|MUL)\b',x))
            nl=sum(1 for _,x in body if 'LDS' in x)
            nloc=sum(1 for _,x in body if 'LDL' in x or 'STL' in x)
            nmov=sum(1 for _,x in body if 'MOV' in x)
            if nf>=20 and len(body)<int(sys.argv[2]): loops.append((len(body),nf,nl,nloc,nmov,hex(tgt),hex(a)))
for l in sorted(loops,key=lambda x:x[0]): print(l)
