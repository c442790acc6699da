import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0])
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name'].split('(')[0][:60]
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1e3
    elif unit=='ms': v*=1e3
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print("%-62s n=%4d total %10.1f us  avg %9.1f us  share %5.1f%%"%(k,v[0],v[1],v[1]/v[0],100*v[1]/tot))
