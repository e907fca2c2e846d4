import json,glob,sys,csv,collections
def bench(pattern):
    for f in sorted(glob.glob(pattern)):
        try: d=json.loads(open(f).read().strip().splitlines()[-1])
        except Exception as e: print(f,'ERR',open(f).read()[-300:]); continue
        r=d['roofline']
        pr=' '.join('%s=%.2f/%d'%(k[:12],v['ms_total'],v['launches']) for k,v in r['profile'].items())
        print("%-28s value %7.1f ms/step %.3f e2e %.2f L/step %.1f whole %.3f | %s"%(f.split('/')[-1],d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']/d['steps'],r['whole_step']['frac'],pr))
def launches(f):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); ui=H.index('Metric Unit')
    agg=collections.defaultdict(lambda:[0,0.0])
    for r in rows[hdr+1:]:
        if len(r)<=vi: continue
        name=r[ki][:90]; v=float(r[vi].replace(',',''))
        if r[ui]=='ns': v/=1e3
        agg[name][0]+=1; agg[name][1]+=v
    tot=sum(v[1] for v in agg.values())
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
        print("%6d launches %10.1f us total %8.2f us avg %5.1f%%  %s"%(v[0],v[1],v[1]/v[0],100*v[1]/tot,k))
if __name__=='__main__':
    if sys.argv[1]=='bench': bench(sys.argv[2])
    else: launches(sys.argv[2])
