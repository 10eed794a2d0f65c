"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ dram__bytes_read/write.sum]) per kernel."""
import csv, re, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter(); rd = collections.defaultdict(float); wr = collections.defaultdict(float)
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']; name = row['Metric Name']
    k = re.sub(r'^void ', '', re.sub(r'\(.*', '', row['Kernel Name'])).replace('mrisr::', '')
    if name == 'gpu__time_duration.sum':
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        tot[k] += v; cnt[k] += 1
    else:
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        (rd if 'read' in name else wr)[k] += v
T = sum(tot.values())
print(f"total {T/1e3:.2f} ms, {sum(cnt.values())} launches, DRAM read {sum(rd.values())/1e9:.2f} GB, write {sum(wr.values())/1e9:.2f} GB")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    extra = f"  dram R {rd[k]/1e6:9.1f} MB W {wr[k]/1e6:9.1f} MB  ({(rd[k]+wr[k])/v/1e6:5.2f} TB/s)" if k in rd else ""
    print(f"{v:10.1f} us {100*v/T:5.1f}%  n={cnt[k]:4d}  {k[:60]:60s}{extra}")
