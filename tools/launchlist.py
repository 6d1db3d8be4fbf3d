import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; data=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
agg=collections.OrderedDict()
for d in data:
    name=re.sub(r'\(.*','',d['Kernel Name']); v=float(d['Metric Value'].replace(',',''))
    unit=d['Metric Unit']
    if unit=='us': v*=1e3
    elif unit=='ms': v*=1e6
    a=agg.setdefault(name,[0,0.0,[]]); a[0]+=1; a[1]+=v; a[2].append(v/1e3)
tot=sum(a[1] for a in agg.values())
print('total us %.1f launches %d'%(tot/1e3,len(data)))
for k,(n,t,l) in sorted(agg.items(), key=lambda x:-x[1][1]):
    print('%-40s n=%4d total=%10.1f us avg=%9.1f us share=%5.1f%%'%(k[:40],n,t/1e3,t/1e3/n,100*t/tot))
    if len(sys.argv)>2 and re.search(sys.argv[2],k): print('     ', ' '.join('%.0f'%x for x in l[:24]))
