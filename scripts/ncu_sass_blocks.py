"""Basic-block level view of an ncu source-page export (`ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv`):
    python scripts/ncu_sass_blocks.py x.csv blocks [n_instructions]   runs of SASS instructions with one execution count: share of the
                                                                      issued warp instructions, active lanes, opcode mix
    python scripts/ncu_sass_blocks.py x.csv stalls [n_instructions]   warp-stall sampling: share of samples per reason
    python scripts/ncu_sass_blocks.py x.csv list                      every instruction
Used for profiles/r2_ncu_stalls_*.md."""
import csv,sys
def f(x):
    try: return float(x.replace(',',''))
    except: return 0.0
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]; data=[r for r in rows[2:] if len(r)>10 and r[0].startswith('0x')]
# there may be two function sections (k_traverse + callee); take until address decreases
iS=h.index("# Samples"); iI=h.index("Instructions Executed"); iT=h.index("Thread Instructions Executed")
totI=sum(f(r[iI]) for r in data); totS=sum(f(r[iS]) for r in data); totT=sum(f(r[iT]) for r in data)
print("n",len(data),"inst",totI,"samples",totS,"avg lanes",totT/totI)
stall_cols=[i for i,k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
mode=sys.argv[2] if len(sys.argv)>2 else "list"
if mode=="list":
    for k,r in enumerate(data):
        ins=f(r[iI]); smp=f(r[iS])
        st=sorted(((f(r[i]),h[i][6:]) for i in stall_cols),reverse=True)[:2]
        print(f"{k:4d} {100*ins/totI:5.2f}%i {100*smp/totS:5.2f}%s L={f(r[iT])/max(ins,1):4.1f} {' '.join(f'{n}:{100*v/max(smp,1):.0f}' for v,n in st if v>0):28s} {r[1].strip()[:90]}")
if mode=="blocks":
    data=data[:int(sys.argv[3])] if len(sys.argv)>3 else data
    totI=sum(f(r[iI]) for r in data); totS=sum(f(r[iS]) for r in data); totT=sum(f(r[iT]) for r in data)
    print("inst",totI,"avg lanes",totT/totI)
    # block = run of instructions with same executed count
    blocks=[];cur=None
    for k,r in enumerate(data):
        ins=f(r[iI])
        if cur is None or abs(ins-cur["ins"])>1e-9*max(ins,1):
            cur={"start":k,"ins":ins,"n":0,"thr":0.0,"smp":0.0,"ops":{}}; blocks.append(cur)
        cur["n"]+=1; cur["thr"]+=f(r[iT]); cur["smp"]+=f(r[iS])
        op=r[1].split()
        op=[o for o in op if not o.startswith('@')][0].split('.')[0]
        cur["ops"][op]=cur["ops"].get(op,0)+1
    for b in blocks:
        share=b["ins"]*b["n"]/totI
        if share>0.003:
            ops=' '.join(f"{k}{v}" for k,v in sorted(b["ops"].items(),key=lambda x:-x[1])[:8])
            print(f"[{b['start']:4d}+{b['n']:3d}] {100*share:5.1f}%i {100*b['smp']/totS:5.1f}%s lanes={b['thr']/max(b['ins']*b['n'],1):4.1f} execs={b['ins']/1e6:7.2f}M  {ops}")
if mode=="stalls":
    n=int(sys.argv[3]) if len(sys.argv)>3 else len(data)
    data=data[:n]
    tot={}
    for r in data:
        for i in stall_cols:
            tot[h[i]]=tot.get(h[i],0)+f(r[i])
    s=sum(tot.values())
    for k,v in sorted(tot.items(),key=lambda x:-x[1]):
        if v/s>0.003: print(f"{k:28s} {100*v/s:5.1f}%")
