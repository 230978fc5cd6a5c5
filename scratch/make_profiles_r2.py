#!/usr/bin/env python
"""Turns the raw ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/:
  launches CSV (--metrics gpu__time_duration.sum)  -> profiles/<tag>_launches_summary.csv
  --set full report (.ncu-rep, n^3 block)          -> profiles/<tag>_ncu_full_n<n>.txt + profiles/ncu_traffic.json
usage: python scratch/make_profiles_r2.py <tag> <launches.csv> <bench.json> <full.ncu-rep> <n of the full capture>"""
import collections, csv, json, re, subprocess, sys

tag, launches, benchjson, rep, nfull = sys.argv[1:6]
nfull = int(nfull)
FP64_PEAK_TDFMA = 17.0      # profiles/r2_dmma_ubench.txt: DFMA, 8 chains, 32 warps/SM


def key_of(name):
    return ("update" if "brick_update" in name else "element_tangent" if "brick_tangent" in name else
            "assemble_A" if "assemble_A" in name else "assemble_B" if "assemble_B" in name else None)


# ---- launch list ----
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name, unit, val = r[4], r[13], float(r[14])
    ms = val / 1e6 if unit in ("nsecond", "ns") else (val / 1e3 if unit in ("usecond", "us") else val)
    per.setdefault(re.sub(r"\(.*", "", name), []).append(ms)
b = json.loads(open(benchjson).read().strip().splitlines()[-1])
km = b["kernel_ms"]
# a step launches the tangent / assembly kernels once per range: shares are taken over per-step sums
per_step = collections.OrderedDict()
steps_seen = max(1, len(per.get(next((k for k in per if "brick_update" in k), ""), [1])))
for k, v in per.items():
    if key_of(k):
        per_step[k] = sum(v) / steps_seen
tot = sum(per_step.values())
with open(f"profiles/{tag}_launches_summary.csv", "w") as f:
    f.write(f"# {tag}: ncu launch list of `bench.py --steps 2 --warmup 3` at n=160 (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
    f.write(f"# live CUDA-event numbers of the same step: update {km['update']:.2f}, assemble_B {km['assemble_B']:.2f}, "
            f"element_tangent {km['element_tangent']:.2f}, assemble_A {km['assemble_A']:.2f} ms; ms_per_step {b['ms_per_step']:.2f} "
            f"(event shares: " + ", ".join(f"{k} {v / (km['update'] + km['assemble_B'] + km['element_tangent'] + km['assemble_A']):.3f}"
                                           for k, v in km.items() if k in ("update", "assemble_B", "element_tangent", "assemble_A")) + ")\n")
    f.write("kernel,launches,mean_ms,ms_per_step,share_of_step\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        mean = sum(v) / len(v)
        f.write(f"{k},{len(v)},{mean:.3f},{per_step[k]:.3f},{per_step[k] / tot:.3f}\n" if k in per_step else f"{k},{len(v)},{mean:.3f},,\n")
print(open(f"profiles/{tag}_launches_summary.csv").read())

# ---- full capture ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]


def num(r, name):
    return float(r[hdr.index(name)]) if name in hdr and r[hdr.index(name)] not in ("", "n/a") else 0.0


def to_bytes(r, name):
    u = units[hdr.index(name)]
    return num(r, name) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)


out, traffic = [], {}
for r in rr[2:]:
    if len(r) < len(hdr):
        continue
    out.append("-----")
    for w in want:
        if w in hdr:
            out.append(f"{w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
    cyc = num(r, 'smsp__cycles_elapsed.avg') or num(r, 'sm__cycles_elapsed.avg')
    fp64 = sum(num(r, f'smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed') for op in ("dfma", "dmul", "dadd")) * cyc
    out.append(f"fp64 lane operations (dfma + dmul + dadd, thread level) = {fp64:.4g}")
    st = sorted(((float(r[hdr.index(h)]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                 for h in stall), reverse=True)[:6]
    out.append('top stalls: ' + ', '.join(f"{n} {v:.2f}" for v, n in st))
    k = key_of(r[hdr.index('Kernel Name')])
    if k and k not in traffic:
        traffic[k] = {"kernel": r[hdr.index('Kernel Name')],
                      "dram_bytes_per_launch": to_bytes(r, 'dram__bytes_read.sum') + to_bytes(r, 'dram__bytes_write.sum'),
                      "fp64_lane_ops_per_launch": fp64,
                      "ms_under_ncu": num(r, 'gpu__time_duration.sum')}
        l1 = num(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')
        if l1 > 80.0:
            traffic[k]["limiter"] = {"unit": "l1_data_pipe", "pct_of_peak": l1,
                                     "note": "ncu l1tex__data_pipe_lsu_wavefronts: shared-memory accumulator + gathers saturate the L1 data pipe"}
open(f"profiles/{tag}_ncu_full_n{nfull}.txt", "w").write(
    f"# {tag}: ncu --set full --clock-control none, bench.py --n {nfull} ({nfull**3} elements); one launch per kernel\n" + "\n".join(out) + "\n")
res = {"source": f"profiles/{tag}_ncu_full_n{nfull}.txt (ncu --set full, one launch per kernel, {nfull}^3 = {nfull**3} elements: dram__bytes_read.sum + "
                 "dram__bytes_write.sum, thread-level dfma + dmul + dadd counts); bench.py scales them by the element count of the run",
       "elements": nfull ** 3, "fp64_peak_tdfma": FP64_PEAK_TDFMA,
       "fp64_peak_source": "profiles/r2_dmma_ubench.txt (DFMA, 8 chains per thread, 32 warps per SM)", "kernels": traffic}
json.dump(res, open("profiles/ncu_traffic.json", "w"), indent=1)
print("\n".join(out))
print(json.dumps(res, indent=1))
