"""Top stall sites of an ncu source-page CSV: python scripts/ncu_hot.py report.ncu-rep [topN] [kernel-index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
# find header line
hi = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[hi:]))))
hdr = rows[0]; body = [r for r in rows[1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stalls}
print("by reason:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
order = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:top]
for i in sorted(order):
    r = body[i]
    why = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*int(r[S])/tot:5.1f}%  {r[ci['Source']][:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
