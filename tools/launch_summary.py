import csv,collections,sys
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    if row.get('Metric Name')=='gpu__time_duration.sum':
        v=float(row['Metric Value'].replace(',',''))
        agg.setdefault(row['Kernel Name'][:70],[]).append(v)
tot=sum(sum(v[-4:]) for v in agg.values())
for k,v in agg.items():
    print('%-72s n=%3d last4 avg %10.1f us  share %5.1f%%'%(k,len(v),sum(v[-4:])/len(v[-4:])/1e3, 100*sum(v[-4:])/tot))
