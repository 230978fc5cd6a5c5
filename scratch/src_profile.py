import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ia=hdr.index('Source'); isamp=hdr.index('Warp Stall Sampling (All Samples)'); iex=hdr.index('Instructions Executed')
data=[r for r in rows[2:] if len(r)>iex and r[isamp].isdigit()]
tot=sum(int(r[isamp]) for r in data)
print("total samples", tot, "n instr", len(data))
step=int(sys.argv[2]) if len(sys.argv)>2 else 100
for i in range(0,len(data),step):
    seg=data[i:i+step]
    s=sum(int(r[isamp]) for r in seg); ex=sum(int(r[iex] or 0) for r in seg)
    nd=sum(1 for r in seg if any(k in r[ia] for k in ('DFMA','DMUL','DADD')))
    nlds=sum(1 for r in seg if 'LDS' in r[ia]); nldg=sum(1 for r in seg if 'LDG' in r[ia]); nst=sum(1 for r in seg if 'STG' in r[ia] or 'STS' in r[ia])
    print("instr %4d: samples %6d (%4.1f%%) exec/instr %8d  fp64 %3d lds %3d ldg %3d st %3d"%(i,s,100*s/tot,ex//max(len(seg),1),nd,nlds,nldg,nst))
top=sorted(data,key=lambda r:-int(r[isamp]))[:12]
for r in top: print(r[isamp], r[ia].strip()[:90])
