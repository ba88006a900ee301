import csv,collections,re,sys
f=sys.argv[1] if len(sys.argv)>1 else '/root/repo/gpurun_out/launches_bf16.csv'
rows=list(csv.reader(open(f)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; data=rows[hi+1:]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict(); tot=0
for r in data:
    if len(r)<=vi: continue
    name=r[ki]; v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1000
    elif r[ui]=='ms': v*=1000
    short=re.sub(r'b2pn::|tc::|simt::|void ','',name); short=re.sub(r'\(.*','',short)[:90]
    agg.setdefault(short,[0,0.0]); agg[short][0]+=1; agg[short][1]+=v; tot+=v
print("total us",round(tot,1),"launches",len(data))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 24]:
    print(f"{t:9.1f} us {n:3d}x {100*t/tot:5.1f}%  {k}")
