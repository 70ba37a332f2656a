import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2] if len(sys.argv) > 2 else '/root/repo/mobilenet_yolo_pytorch_b200/csrc/decode_nms.cuh').read().split('\n')
# function ranges by scanning for known function names
marks=[]
for i,l in enumerate(src,1):
    m=re.match(r'(?:template.*\n)?__device__.*?\b(\w+)\(', l) or re.match(r'__global__.*?\b(\w+)\(', l)
    if m and not l.startswith(' '): marks.append((i,m.group(1)))
def fn(ln):
    name='?'
    for i,n in marks:
        if i<=ln: name=n
        else: break
    return name
fname=""; hdr=None; out=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if len(r)>5 and r[0]=="Line No": hdr={n:k for k,n in enumerate(r)}; continue
    if hdr is None or len(r)<10 or r[2]!="-": continue
    try:
        inst=int(float(r[hdr["Instructions Executed"]] or 0)); samp=int(float(r[hdr["# Samples"]] or 0))
    except ValueError: continue
    out.append((fname,int(r[0]),inst,samp))
ti=sum(o[2] for o in out); ts=sum(o[3] for o in out)
agg={}
for f,ln,inst,samp in out:
    key = fn(ln) if f=="decode_nms.cuh" else f
    d=agg.setdefault(key,[0,0]); d[0]+=inst; d[1]+=samp
N=256
for k,(i,s) in sorted(agg.items(), key=lambda x:-x[1][0]):
    if i/ti>0.002 or s/ts>0.005: print(f"{k:28s} inst {100*i/ti:5.1f}% ({i/N/1000:6.1f}k/img)  samples {100*s/ts:5.1f}%")
print("total", ti/N/1000, "k/img", "samples", ts)
