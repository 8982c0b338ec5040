import sys,collections,re
sys.path.insert(0,'tools')
import ncu_lines as nl
rep,obj,kern,mang,warps=sys.argv[1],sys.argv[2],sys.argv[3],sys.argv[4],int(sys.argv[5])
blk=nl.sass_page(rep,kern,0)
hdr=blk['hdr']; ci={n:hdr.index(n) for n in ("Instructions Executed","Thread Instructions Executed","# Samples","Source")}
li=nl.line_info(obj,mang,'skyjo_core.cuh')
rows=blk['rows']
print(len(rows),len(li))
agg=collections.defaultdict(lambda:[0,0,0])
tot=0
for k in range(len(rows)):
    r=rows[k]; inst=int(r[ci["Instructions Executed"]]); t=int(r[ci["Thread Instructions Executed"]])
    tot+=inst
    if inst and t/inst<4:
        g=agg[li[k][1]]; g[0]+=inst; g[1]+=t; g[2]+=1
print("low-lane lines (per warp-step):")
for key,g in sorted(agg.items(), key=lambda kv:-kv[1][0])[:45]:
    print(f"{key[0]}:{key[1]:4d} sass {g[2]:4d} inst/warp-step {g[0]/warps:7.2f} lanes {g[1]/g[0]:.1f}")
print("sum", sum(g[0] for g in agg.values())/warps, "total", tot/warps)
