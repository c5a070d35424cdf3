# aggregate ncu per-instruction samples / executed counts by source line (nvdisasm -g line info, matched by instruction index)
import re,csv,collections,sys
sass, func, csvf, nw = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lines=open(sass).read().split('\n')
start=None
for i,l in enumerate(lines):
    if l.startswith('.text.'+func) and l.endswith(':'): start=i
cur=None; seq=[]
for l in lines[start+1:]:
    if l.startswith('//-------') : break
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*);',l)
    if m: seq.append((m.group(2).strip(),cur))
rows=list(csv.reader(open(csvf)))
hdr=rows[1]; ix=hdr.index("Instructions Executed"); ismp=hdr.index("# Samples")
body=rows[2:]
assert len(body)==len(seq),(len(body),len(seq))
ex=collections.Counter(); sm=collections.Counter()
for k in range(len(seq)):
    ex[seq[k][1]]+=int(body[k][ix]); sm[seq[k][1]]+=int(body[k][ismp])
ts=sum(sm.values()); te=sum(ex.values())
print("total samples",ts,"instr/warp",te/nw)
src={}
for key,v in sm.most_common(int(sys.argv[5]) if len(sys.argv)>5 else 40):
    f,l=key if key else ('?',0)
    if f not in src:
        try: src[f]=open('/root/repo/rabi_b200/csrc/'+f).read().split('\n')
        except Exception: src[f]=[]
    text=src[f][l-1].strip()[:90] if 0<l<=len(src[f]) else ''
    print(f"{f}:{l:4d} time {100*v/ts:5.1f}%  instr {100*ex[key]/te:5.1f}%  | {text}")
