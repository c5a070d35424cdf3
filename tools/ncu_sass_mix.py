import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1]))); nw=float(sys.argv[2])
hdr=rows[1]; ix=hdr.index("Instructions Executed"); isrc=hdr.index("Source"); ismp=hdr.index("# Samples"); iw=hdr.index("L1 Wavefronts Shared")
cnt=collections.Counter(); smp=collections.Counter(); wf=collections.Counter()
def opof(s):
    t=s.split(); op=t[1] if t[0].startswith('@') else t[0]
    return op
for r in rows[2:]:
    op=opof(r[isrc]); b=op.split('.')[0]
    if b in('LDS','STS','ATOMS'): b=op
    cnt[b]+=int(r[ix]); smp[b]+=int(r[ismp]); wf[b]+=int(r[iw] or 0)
tot=sum(cnt.values()); ts=sum(smp.values())
for k,v in cnt.most_common(28): print(f"{k:16s} {v/nw:10.0f} {100*v/tot:5.1f}%  samples {100*smp[k]/ts:5.1f}%  wavefronts/inst {wf[k]/max(v,1):.2f}")
print(tot/nw)
segs=[];cur=None
for i,r in enumerate(rows[2:]):
    c=int(int(r[ix])/nw)
    if cur and abs(cur[2]-c)<=max(2,0.02*c): cur[1]=i; cur[3].append(opof(r[isrc]).split('.')[0])
    else:
        cur=[i,i,c,[opof(r[isrc]).split('.')[0]]]; segs.append(cur)
out=[((s[1]-s[0]+1)*s[2],s) for s in segs]
for w,s in sorted(out,key=lambda t:-t[0])[:int(sys.argv[3]) if len(sys.argv)>3 else 16]:
    c=collections.Counter(s[3])
    print(f"lines {s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:4d} exec/warp={s[2]:6d} weight={w:8d} ({100*w/(tot/nw):4.1f}%)", dict(c.most_common(8)))
