import csv, collections, subprocess, sys
rep=sys.argv[1]
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
import io
rows=list(csv.reader(io.StringIO(src)))
hdr=rows[1]; data=rows[2:]
iA,iS,iN,iE=(hdr.index(k) for k in ('Address','Source','# Samples','Instructions Executed'))
base=int(data[0][iA],16)
tot=sum(int(r[iN]) for r in data)
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
segs=[]; cur=None
for r in data:
    a=int(r[iA],16)-base; e=int(r[iE]); n=int(r[iN])
    if cur is None or abs(e-cur['e'])>0.02*max(e,cur['e'],1):
        cur={'a0':a,'e':e,'n':0,'cnt':0,'inst':0,'st':collections.Counter()}; segs.append(cur)
    cur['n']+=n; cur['cnt']+=1; cur['a1']=a; cur['inst']+=e
    for i in stall_cols: cur['st'][hdr[i][6:]]+=int(r[i])
print("total samples",tot,"total inst %.1fM"%(sum(s['inst'] for s in segs)/1e6))
for s in segs:
    if s['n']>0.004*tot or s['inst']>3e6:
        top=", ".join(f"{k}:{100*v/max(1,s['n']):.0f}" for k,v in s['st'].most_common(5))
        print(f"{s['a0']:6x}-{s['a1']:6x} instrs {s['cnt']:5d} exec/instr {s['e']:9d} inst {s['inst']/1e6:7.1f}M samples {100*s['n']/tot:5.1f}%  [{top}]")
