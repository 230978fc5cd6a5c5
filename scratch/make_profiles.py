#!/usr/bin/env python
"""Turns the raw ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/:
  launches CSV (--metrics gpu__time_duration.sum)  -> profiles/<tag>_launches_summary.csv
  --set full report (.ncu-rep, n^3 block)          -> profiles/<tag>_ncu_full.txt + profiles/ncu_traffic.json
usage: python scratch/make_profiles.py <tag> <launches.csv> <bench.json> <full.ncu-rep> <n of the full capture>"""
import csv, json, subprocess, sys, collections, re

tag, launches, benchjson, rep, nfull = sys.argv[1:6]
nfull = int(nfull)
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name, unit, val = r[4], r[13], float(r[14])
    ms = val / 1e6 if unit in ("nsecond", "ns") else (val / 1e3 if unit in ("usecond", "us") else val)
    per.setdefault(re.sub(r"\(.*", "", name), []).append(ms)
b = json.load(open(benchjson))
step_kernels = [k for k in per if any(s in k for s in ("brick_update", "assemble_B", "brick_tangent", "assemble_A"))]
tot = sum(sum(per[k]) / len(per[k]) for k in step_kernels)
with open(f"profiles/{tag}_launches_summary.csv", "w") as f:
    f.write(f"# {tag}: ncu launch list of `bench.py --steps 2 --warmup 3` at n=160 (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
    km = b["kernel_ms"]
    f.write(f"# live CUDA-event numbers of the same step ({benchjson}): update {km['update']:.2f}, assemble_B {km['assemble_B']:.2f}, "
            f"element_tangent {km['element_tangent']:.2f}, assemble_A {km['assemble_A']:.2f} ms; ms_per_step {b['ms_per_step']:.2f}\n")
    f.write("kernel,launches,mean_ms,share_of_step\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
        mean = sum(v) / len(v)
        f.write(f"{k},{len(v)},{mean:.3f},{mean / tot:.3f}\n" if k in step_kernels else f"{k},{len(v)},{mean:.3f},\n")
ev = {"update": km["update"], "assemble_B": km["assemble_B"], "element_tangent": km["element_tangent"], "assemble_A": km["assemble_A"]}
print("event shares", {k: round(v / sum(ev.values()), 3) for k, v in ev.items()})
print(open(f"profiles/{tag}_launches_summary.csv").read())

out = subprocess.run([sys.executable, "scratch/ncu_summary.py", rep], capture_output=True, text=True).stdout
open(f"profiles/{tag}_ncu_full_n{nfull}.txt", "w").write(
    f"# {tag}: ncu --set full --clock-control none, bench.py --n {nfull} ({nfull**3} elements); one launch per kernel\n" + out)
traffic = {}
cur = None
for line in out.splitlines():
    if line.startswith("Kernel Name"):
        cur = line.split("=", 1)[1].strip()
        key = ("update" if "brick_update" in cur else "element_tangent" if "brick_tangent" in cur else
               "assemble_A" if "assemble_A" in cur else "assemble_B" if "assemble_B" in cur else None)
        if key and key not in traffic:
            traffic[key] = {"kernel": cur, "read": 0.0, "write": 0.0}
        else:
            key = None if key in traffic and traffic[key].get("done") else key
        curkey = key
    elif line.startswith("dram__bytes_read.sum") and curkey and "done" not in traffic[curkey]:
        v, u = line.split("=")[1].split()[:2]; traffic[curkey]["read"] = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
    elif line.startswith("dram__bytes_write.sum") and curkey and "done" not in traffic[curkey]:
        v, u = line.split("=")[1].split()[:2]; traffic[curkey]["write"] = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        traffic[curkey]["done"] = True
res = {"source": f"profiles/{tag}_ncu_full_n{nfull}.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch, "
                 f"{nfull}^3 = {nfull**3} elements); bench.py scales it by the element count of the run",
       "elements": nfull ** 3, "kernels": {k: {"dram_bytes_per_launch": v["read"] + v["write"], "kernel": v["kernel"]} for k, v in traffic.items()}}
json.dump(res, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(res, indent=1))
