"""profiles/r2_traffic_b32.json from an ncu launch list of one batch-32 denoising step
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ... scripts/profile_step.py 32`):
per kernel class, launches and DRAM bytes (read + write).  bench.py reports `roofline.traffic` from this file (per launch,
scaled by batch / 32) -- measured by ncu this round, never a constant in the source.
    python scripts/traffic_from_ncu.py profiles/r2_launches_step_b32.csv <git commit of the profiled tree> [output json]"""
import collections, csv, json, re, sys
src, commit = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "unknown")
lines = [l for l in open(src) if not l.startswith("==")]
cls_of = [(r"gemm_tcgen05_kernel", "gemm"), (r"attention_tcgen05_split_kernel<40", "attn_self_d40"), (r"attention_ctx_kernel", "attn_cross"),
          (r"attention", "attn_self"), (r"groupnorm", "groupnorm"), (r"layernorm", "layernorm"), (r"sched_step", "sched_step")]
acc = collections.defaultdict(lambda: {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    c = next((c for pat, c in cls_of if re.search(pat, name)), "other")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]; m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        acc[c]["launches"] += 1
        acc[c]["us"] += v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    else:
        acc[c]["dram_bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = {"batch": 32, "source": f"{src} (ncu launch list of one eager batch-32 step, commit {commit}): dram__bytes_read.sum + dram__bytes_write.sum per kernel class / launches",
       "classes": dict(acc)}
json.dump(out, open(sys.argv[3] if len(sys.argv) > 3 else "profiles/r2_traffic_b32.json", "w"), indent=1)
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k:14s} n={v['launches']:4d} {v['us']:9.1f} us  {v['dram_bytes']/1e6:9.1f} MB")
