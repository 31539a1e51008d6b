import csv,sys,collections,re
rows=list(csv.reader(open(sys.argv[1], errors='ignore')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); 
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    name=re.sub(r'\(.*','',r[kn]); 
    try: v=float(r[mv].replace(',',''))
    except: continue
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
unit=rows[hi+1][h.index('Metric Unit')] if 'Metric Unit' in h else ''
print('total',tot,unit)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]:
    print(f"{v[1]/tot*100:5.1f}%  n={v[0]:4d}  avg={v[1]/v[0]:10.1f}  {k[:90]}")
