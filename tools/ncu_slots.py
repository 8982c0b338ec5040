"""Development aid: per-region / low-lane breakdown of one captured launch (warp-instructions per warp)."""
import sys,collections
sys.path.insert(0,'tools')
import ncu_lines as nl
rep,obj,kern,mang,W=sys.argv[1],sys.argv[2],sys.argv[3],sys.argv[4],int(sys.argv[5])
blk=nl.sass_page(rep,kern,0)
hdr=blk['hdr']; ci={n:hdr.index(n) for n in ("Instructions Executed","Thread Instructions Executed","# Samples","Source")}
li=nl.line_info(obj,mang,'skyjo_step.cuh')
li2=nl.line_info(obj,mang,'skyjo_core.cuh')
rows=blk['rows']
assert len(rows)==len(li),(len(rows),len(li))
tot=0; low=collections.defaultdict(int); reg=collections.defaultdict(int); b=collections.defaultdict(int); mid=collections.defaultdict(int)
for k,r in enumerate(rows):
    i=int(r[ci["Instructions Executed"]]); t=int(r[ci["Thread Instructions Executed"]]); tot+=i
    reg[li[k][1]]+=i
    if i:
        l=t/i
        b["<4" if l<4 else "<12" if l<12 else "<20" if l<20 else "<28" if l<28 else ">=28"]+=i
        if l<4: low[li2[k][1]]+=i
        elif l<28: mid[li2[k][1]]+=i
print(rep,"total/warp",round(tot/W,1), {k:round(v/W,1) for k,v in b.items()})
print(" regions:",[(f"{k[0].replace('skyjo_','')}:{k[1]}",round(v/W,1)) for k,v in sorted(reg.items(),key=lambda kv:-kv[1])[:10]])
print(" low-lane:",[(f"{k[0].replace('skyjo_','')}:{k[1]}",round(v/W,1)) for k,v in sorted(low.items(),key=lambda kv:-kv[1])[:16]])
print(" mid-lane:",[(f"{k[0].replace('skyjo_','')}:{k[1]}",round(v/W,1)) for k,v in sorted(mid.items(),key=lambda kv:-kv[1])[:16]])
