import csv,collections,sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0])
n=0
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    k=row['Kernel Name'].replace('void ','').replace('jmt::','')[:50]
    v=float(row['Metric Value'].replace(',',''))
    u=row['Metric Unit']
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    agg[k][0]+=1; agg[k][1]+=v; n+=1
tot=sum(a[1] for a in agg.values())
print(n,'launches',round(tot,1),'us total')
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:50s} {a[0]:5d} {a[1]:10.1f} us {100*a[1]/tot:5.1f}%  avg {a[1]/a[0]:7.1f}")
