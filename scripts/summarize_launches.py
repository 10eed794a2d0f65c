import csv, re, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
tot=collections.defaultdict(float); cnt=collections.Counter()
for row in csv.DictReader(lines):
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    k=re.sub(r'^void ','',re.sub(r'\(.*','',row['Kernel Name'])).replace('mrisr::','')
    tot[k]+=v; cnt[k]+=1
T=sum(tot.values())
print(f"total {T/1e3:.2f} ms, {sum(cnt.values())} launches")
for k,v in sorted(tot.items(), key=lambda x:-x[1]):
    print(f"{v:10.1f} us {100*v/T:5.1f}%  n={cnt[k]:4d}  {k[:80]}")
