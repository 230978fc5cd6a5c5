#!/usr/bin/env python
"""bench.py -- the reference's headline hot path on B200.

Metric (BASELINE.json): Gauss-point state updates/s through one pass of the hot path
  step = Domain::update (state determination) + formUnbalance + formTangent
on the 3D stdBrick / J2Plasticity block (configs[2], 160^3 = 4.096M elements, 32.8M Gauss
points, 12.4M equations, 0.98G non-zeros), synthetic mesh and displacement field.

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path)

One JSON line on stdout (rank 0).  Timing: CUDA events on the stream the kernels are
launched on, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "gauss_point_state_updates_per_s (update + formUnbalance + formTangent pass)"
UNIT = "GP updates/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """samples taken between t0 and t1 (the timed region, with 0.1 s of margin); a region shorter than the
        sampling period falls back on the sample closest to it"""
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        rows = [(t, r) for t, r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if t0 is not None and rows:
            inside = [(t, r) for t, r in rows if t0 - 0.1 <= t <= t1 + 0.1]
            rows = inside or [min(rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t1)))]
        rows = [r for _, r in rows]
        sm = [float(r[1]) for r in rows]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def workload_spec(n, workload="brick"):
    """brick: n^3 J2 stdBrick block (BASELINE configs[2], the headline workload);
    quad: n x n plane-strain ElasticIsotropic FourNodeQuad mesh (configs[1]);
    frame: 2D RC frame of forceBeamColumn elements with Steel02/Concrete02 fibre sections, n bays x n storeys
           (the 2D counterpart of configs[3]);
    frame3d: 3D RC space frame, forceBeamColumn (ForceBeamColumn3d) + FiberSection3d (configs[3])"""
    from modelspec import ELASTIC, J2_STEEL, brick_block, frame2d, frame3d, quad_plane, soil_structure_block
    if workload == "soil":      # configs[4]: J2 soil block with an elastic footing + pier embedded at the top (two element batches)
        return soil_structure_block(n, n, n)
    if workload == "quad":
        return quad_plane(n, n, mat=(ELASTIC[0], [1000.0, 0.25, 0.0]), lx=float(n), ly=float(n))
    if workload == "frame":
        return frame2d(n, n, 2)
    if workload == "frame3d":   # n x n bays, 3.8 n storeys, 2 elements per member: n = 20 -> 194,712 elements
        return frame3d(n, n, max(1, round(3.8 * n)), ndiv=2)
    return brick_block(n, n, n, mat=J2_STEEL)


def displacement_field_for(spec, crd, workload):
    if workload in ("brick", "soil"):
        return displacement_field(crd)
    if workload == "quad":
        x, y = crd[:, 0], crd[:, 1]
        u = np.empty_like(crd)
        u[:, 0] = 1e-3 * (y * y / (1.0 + y.max()) + 0.3 * np.sin(0.07 * x) * y)
        u[:, 1] = 1e-3 * (0.5 * x * y / (1.0 + x.max()))
        return u
    if workload == "frame3d":                            # biaxial sway with matching joint rotations + a floor twist
        H = crd[:, 2].max()
        z = crd[:, 2] / H
        xc, yc = crd[:, 0] - crd[:, 0].mean(), crd[:, 1] - crd[:, 1].mean()
        a = 0.002 * H
        th = 1e-4 * z
        u = np.zeros((len(crd), 6))
        u[:, 0] = a * z ** 1.5 - th * yc; u[:, 1] = 0.6 * a * z ** 1.5 + th * xc; u[:, 2] = -0.01 * z
        u[:, 3] = -0.6 * 1.5 * a * z ** 0.5 / H; u[:, 4] = 1.5 * a * z ** 0.5 / H; u[:, 5] = th
        return u
    y = crd[:, 1] / crd[:, 1].max()                      # frame: sway with matching joint rotations
    u = np.zeros((len(crd), 3))
    a = 0.004 * crd[:, 1].max()
    u[:, 0] = a * y ** 1.5; u[:, 1] = -0.01 * y; u[:, 2] = -1.5 * a * y ** 0.5 / crd[:, 1].max()
    return u


def displacement_field(crd, amp=4e-3):
    """smooth synthetic trial displacement: shear + bending + a ripple; strains ~0.1-0.6 %, i.e. a
    mix of elastic and yielded Gauss points for sig0/(2G) ~ 0.16 %"""
    x, y, z = crd[:, 0], crd[:, 1], crd[:, 2]
    u = np.empty_like(crd)
    u[:, 0] = amp * (z * z + 0.3 * np.sin(7.0 * y) * z)
    u[:, 1] = amp * (0.5 * z * x + 0.2 * np.sin(5.0 * x) * z)
    u[:, 2] = amp * (-0.4 * z + 0.3 * x * z * y)
    return u


# ---------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------
def _cpu_worker(args):
    """one process = one core: the reference's own classes (oracle/_ref) when built, else the
    oracle port; times `steps` passes of update + formUnbalance + formTangent on an m^3 sample"""
    m, steps, warmup, use_ref, seed = args
    from modelspec import J2_STEEL, OracleBackend, RefBackend, brick_block
    spec = brick_block(m, m, m, mat=J2_STEEL)
    B = (RefBackend if use_ref else OracleBackend)(spec, 0, 0)
    u = displacement_field(spec.crd)
    u[B.ids() < 0] = 0.0
    B.apply_load(1.0)
    for _ in range(warmup):
        B.set_trial_disp(u); B.form_unbalance(); B.form_tangent()
    t0 = time.perf_counter()
    for _ in range(steps):
        B.set_trial_disp(u)      # Node::setTrialDisp + Domain::update (state determination)
        B.form_unbalance()
        B.form_tangent()
    dt = time.perf_counter() - t0
    return spec.ne * 8 * steps, dt


def cpu_arm(steps, warmup, procs, m):
    from modelspec import have_ref
    use_ref = have_ref()
    if procs == 1:
        res = [_cpu_worker((m, steps, warmup, use_ref, 0))]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, [(m, steps, warmup, use_ref, i) for i in range(procs)])
    gp = sum(r[0] for r in res)
    dt = max(r[1] for r in res)
    return {"value": gp / dt, "unit": UNIT, "cores": procs, "kind": "reference" if use_ref else "port",
            "sample": f"{procs} x ({m}^3 = {m ** 3} stdBrick/J2 elements, {steps} passes of update+formUnbalance+formTangent"
                      f", SparseGenCol addA) ; {dt:.1f} s wall"}, dt / steps


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = max(1, min(os.cpu_count() or 1, 64))
    cb, step_s = cpu_arm(max(1, a.steps), max(1, min(a.warmup, 1)), procs, a.cpu_sample)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a.n), "sample_per_core": f"{a.cpu_sample}^3 elements",
                       "note": "reference CPU path (sequential OpenSees classes), one independent sample per host core"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(n, workload="brick"):
    if workload == "quad":
        return f"2D FourNodeQuad plane-strain ElasticIsotropic mesh {n}x{n} = {n * n} elements (BASELINE configs[1])"
    if workload == "frame":
        return f"2D RC frame {n} bays x {n} storeys, forceBeamColumn + fibre sections Steel02/Concrete02, Newmark terms (2D counterpart of BASELINE configs[3])"
    if workload == "soil":
        return (f"soil-structure block {n}x{n}x{n} = {n ** 3} stdBrick elements: J2Plasticity soil + ElasticIsotropic footing and pier, two element "
                "batches (BASELINE configs[4]; its quads cannot share a Domain with bricks here: one ndf per model)")
    if workload == "frame3d":
        return (f"3D RC space frame {n}x{n} bays x {max(1, round(3.8 * n))} storeys, forceBeamColumn (ForceBeamColumn3d) + FiberSection3d "
                "Steel02/Concrete02, transient Newmark terms (BASELINE configs[3])")
    return f"3D stdBrick J2Plasticity block {n}x{n}x{n} = {n ** 3} elements (BASELINE configs[2])"


# ---------------------------------------------------------------------------------------
# N >= 2 self-test (the driver's GPU test box has one GPU: this is the multi-GPU parity check it can run)
# ---------------------------------------------------------------------------------------
def selftest_main(a):
    """torchrun -N: every rank sets up its partition of a small distorted J2 brick block, the interface rows and
    residuals cross NCCL, and rank 0 -- which also runs the whole model on its GPU -- checks that every rank's owned
    rows of A and B equal the single-GPU rows BIT FOR BIT.  Prints one JSON line {"selftest": "ok", ...}."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    import xara_b200 as xb
    from modelspec import J2_STEEL, brick_block
    torch.cuda.set_device(local)
    if world < 2:
        raise SystemExit("--selftest needs torchrun with at least 2 ranks")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = []
    # the third case is large enough (>= 65536 elements per rank at N = 2) for the two-stream formTangent, where the
    # interface exchange runs beside the assembly of the interior nodes
    for numberer, soe, dims in ((1, 0, (12, 10, 14)), (0, 1, (12, 10, 14)), (0, 0, (56, 52, 26 * world))):
        mk = lambda: brick_block(*dims, mat=J2_STEEL, distort=0.2, seed=3)
        spec = mk()
        D = xb.DeviceModel.from_spec(spec, numberer, soe, world, rank).to_device(local)
        box = [xb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, 0)
        D.comm_init(box[0])
        ug = np.random.default_rng(1).normal(0, 3e-3, (spec.nn, 3))
        for s_ in range(2):
            D.set_trial_disp((s_ + 1) * ug[D.node_tags() - 1]); D.update(); D.apply_load(0.7)
            A, B = D.form_tangent(), D.form_unbalance()
            if s_ == 0:
                D.commit()
        rows = D.row_eqns()
        gathered = [None] * world
        dist.gather_object((rows, A, B), gathered if rank == 0 else None, 0)
        if rank == 0:
            G = xb.DeviceModel.from_spec(mk(), numberer, soe).to_device(local)
            for s_ in range(2):
                G.set_trial_disp((s_ + 1) * ug); G.update(); G.apply_load(0.7)
                Ag, Bg = G.form_tangent(), G.form_unbalance()
                if s_ == 0:
                    G.commit()
            gptr, _ = G.pattern()
            seen = 0
            for rws, Ar, Br in gathered:
                ok = np.array_equal(Br, Bg[rws]) and np.array_equal(Ar, np.concatenate([Ag[gptr[q]:gptr[q + 1]] for q in rws]))
                if not ok:
                    raise SystemExit("selftest FAILED: a rank's rows differ from the single-GPU rows")
                seen += len(rws)
            if seen != G.neq:
                raise SystemExit("selftest FAILED: the ranks' rows do not cover the system")
            res.append({"numberer": numberer, "soe": soe, "elements": int(G.ne), "equations": int(G.neq), "nnz": int(G.nnz)})
            del G
        dist.barrier()
    if rank == 0:
        print(json.dumps({"selftest": "ok", "n_gpus": world, "what": "owned rows of A and B of every rank bitwise equal to the "
                          "single-GPU rows (stdBrick/J2 blocks, two load steps with a commit, NCCL interface exchange; the last case "
                          "runs the two-stream formTangent)",
                          "cases": res}), flush=True)
    dist.destroy_process_group()


def load_invariants():
    p = os.path.join(ROOT, "tests", "golden", "bench_invariants.json")
    return json.load(open(p)) if os.path.exists(p) else {}


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def ours_main(a):
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version
    # there) are pointed at stderr, the JSON goes out through a private copy of the original fd
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and a.gpus > 1:
        raise SystemExit("launch N > 1 with torch.distributed.run (see module docstring)")
    if world > 1:   # the host-side set-up is OpenMP-parallel: share the cores between the ranks
        # (torchrun exports OMP_NUM_THREADS=1, which would serialise the set-up; must be set before
        # libxara_b200.so brings libgomp in)
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or world) // world))

    import torch
    import torch.distributed as dist

    import xara_b200 as xb

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = a.n
    t0 = time.time()
    spec = workload_spec(n, a.workload)
    t_mesh = time.time() - t0
    t0 = time.time()
    is_frame = a.workload in ("frame", "frame3d")
    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.opt}   # xb_set_option, before the set-up: results do not depend on them
    if is_frame:
        # configs[3] is a transient Newmark run: nodal masses (translations), average acceleration, dt = 0.02:
        # formTangent gives c1 K + c3 M, formUnbalance P - M a - R
        D = xb.DeviceModel.from_spec(spec, setup=False, options=opts)
        mass = np.zeros((spec.nn, spec.ndf)); mass[:, :spec.ndm] = 0.05
        D.set_mass(spec.node_tags, mass)
        D.setup(xb.NUMBERER_PLAIN, xb.SOE_SPARSE_GEN_COL, world, rank)
    else:
        part, part_info = None, None
        if a.partition == "metis" and world > 1:
            # DomainPartitioner's way: METIS_PartGraphKway (the reference's own OTHER/METIS, compiled by oracle/ref_build.mk)
            # on Domain::buildEleGraph's element graph, computed on rank 0 and handed to every rank as part[e]
            from modelspec import have_metis, metis_partition
            box = [None]
            if rank == 0:
                if not have_metis():
                    raise SystemExit("--partition metis needs oracle/_ref/libmetis_ref.so (python -c 'import __graft_entry__ as g; g.build()')")
                tm = time.time(); box[0] = metis_partition(spec, world, fast=True); part_info = {"metis_s": time.time() - tm}
            dist.broadcast_object_list(box, 0)
            part = box[0]
        D = xb.DeviceModel.from_spec(spec, xb.NUMBERER_PLAIN, xb.SOE_SPARSE_GEN_COL, world, rank, part, options=opts)
    t_setup = time.time() - t0
    stream = torch.cuda.Stream()          # a real (non-default) stream: the kernels and the events share it
    torch.cuda.set_stream(stream)
    t0 = time.time()
    D.to_device(local, stream=stream.cuda_stream)
    if world > 1:
        box = [xb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, 0)
        D.comm_init(box[0])
    t_upload = time.time() - t0
    # clocks / throttle reasons: ONE nvidia-smi loop (rank 0's GPU), started well before the warm-up so that its NVML
    # set-up is over when the timed region begins (eight of them starting inside a 20 ms region showed as jitter)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ids = D.ids()
    u = displacement_field_for(spec, spec.crd[D.node_tags() - 1], a.workload); u[ids < 0] = 0.0
    nip = {"brick": 8, "soil": 8, "quad": 4, "frame": 5, "frame3d": 4}[a.workload]
    ngp_global, ne_global = spec.ne * nip, spec.ne
    del spec
    if is_frame:
        gamma, beta, dt = 0.5, 0.25, 0.02
        D.set_transient(1.0, gamma / (beta * dt), 1.0 / (beta * dt * dt))
    D.set_trial_disp(u); D.apply_load(1.0); D.synchronize()
    u_alt = (u, 1.01 * u)
    flip = [0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if is_frame:   # a force-based beam returns at once on a zero increment: alternate two trial fields
            flip[0] ^= 1
            D.set_trial_disp(u_alt[flip[0]])
        D.update(); D.form_unbalance(host=False); D.form_tangent(host=False)

    for _ in range(max(a.warmup, 3)):
        step()
    D.synchronize()

    # ---- the timed region: K passes of the hot path through the public calls ----
    l0 = D.launch_count()
    barrier()
    t_wall0 = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for k in range(a.steps):
        step()
    end.record(stream)
    barrier()
    D.synchronize()
    t_wall1 = time.time()
    total_ms = start.elapsed_time(end)
    launches = D.launch_count() - l0
    clk = clocks.stop(t_wall0, t_wall1)

    # ---- the same K passes again with an event between the phases (formTangent un-pipelined here,
    # so that each kernel's own duration is seen): per-kernel times for the roofline ----
    names = ["update", "element_resid", "exchange_B", "assemble_B", "element_tangent", "exchange_A", "assemble_A"]
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(8)] for _ in range(a.steps)]
    barrier()
    for k in range(a.steps):
        e = ev[k]
        if is_frame:
            flip[0] ^= 1
            D.set_trial_disp(u_alt[flip[0]])
        e[0].record(stream); D.update()
        e[1].record(stream); D.form_element_resids()
        e[2].record(stream); D.exchange(1)
        e[3].record(stream); D.assemble_unbalance()
        e[4].record(stream); D.form_element_tangents()
        e[5].record(stream); D.exchange(0)
        e[6].record(stream); D.assemble_tangent()
        e[7].record(stream)
    barrier()
    D.synchronize()
    # xb_form_tangent as the timed region calls it (tiled element -> assembly pipeline on large brick models)
    ft = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    for k in range(a.steps):
        ft[k].record(stream); D.form_tangent(host=False)
    ft[a.steps].record(stream)
    barrier()
    D.synchronize()
    ft_call_ms = float(np.mean([ft[k].elapsed_time(ft[k + 1]) for k in range(a.steps)]))
    ms = {nm: float(np.mean([ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(a.steps)])) for i, nm in enumerate(names)}
    if world > 1:   # time on the device, MAX over ranks
        t = torch.tensor([total_ms] + [ms[nm] for nm in names] + [float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0]); ms_max = {nm: float(t[1 + i]) for i, nm in enumerate(names)}
        tl = torch.tensor([float(launches)], device="cuda", dtype=torch.float64); dist.all_reduce(tl)
        launches_all = int(tl[0])
    else:
        ms_max, launches_all = ms, launches
    ms_per_step = total_ms / a.steps
    value = ngp_global / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches), live numbers ----
    peak, peak_src = peaks()
    which = {"update": 0, "assemble_B": 1, "element_tangent": 3, "assemble_A": 4}
    dom = max(which, key=lambda k: ms[k])
    alg = D.algorithmic_bytes(which[dom])
    achieved = alg / (ms[dom] * 1e-3) / 1e9
    path_alg = D.algorithmic_bytes(0) + D.algorithmic_bytes(1) + D.algorithmic_bytes(2)
    # per-kernel counters from the committed `ncu --set full` capture (profiles/ncu_traffic.json, taken on a smaller
    # block of the same workload), scaled to this rank's element count: DRAM bytes and FP64 lane operations per launch
    tj = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if a.workload == "brick" and os.path.exists(tpath):
        tj = json.load(open(tpath))
    scale_e = float(D.ne) / tj["elements"] if tj else 0.0
    kern = {}
    for nm_, wi in which.items():
        kb = D.algorithmic_bytes(wi)
        rec = {"ms": ms[nm_], "algorithmic_GB": kb / 1e9, "hbm_frac": kb / (ms[nm_] * 1e-3) / 1e9 / peak if ms[nm_] > 0 else None}
        tk = tj.get("kernels", {}).get(nm_, {})
        if "dram_bytes_per_launch" in tk:
            rec["dram_traffic_GB"] = tk["dram_bytes_per_launch"] * scale_e / 1e9
        if "fp64_lane_ops_per_launch" in tk and tj.get("fp64_peak_tdfma"):
            ops = tk["fp64_lane_ops_per_launch"] * scale_e
            rec["fp64"] = {"lane_ops": ops, "achieved_tdfma": ops / (ms[nm_] * 1e-3) / 1e12, "peak_tdfma": tj["fp64_peak_tdfma"],
                           "frac": ops / (ms[nm_] * 1e-3) / 1e12 / tj["fp64_peak_tdfma"]}
        if "limiter" in tk:
            rec["limiter"] = tk["limiter"]
        kern[nm_] = rec
    traffic = kern[dom].get("dram_traffic_GB")
    # the floor of the step: the compulsory bytes at the HBM peak against the FP64 work of the kernels that have any
    # at the measured DFMA peak -- the larger of the two bounds the path from below
    hbm_floor_ms = path_alg / (peak * 1e9) * 1e3
    fp64_floor_ms = (sum(k["fp64"]["lane_ops"] for k in kern.values() if "fp64" in k) / (tj["fp64_peak_tdfma"] * 1e12) * 1e3
                     if tj.get("fp64_peak_tdfma") else None)
    dom_fp64 = kern[dom].get("fp64", {}).get("frac")
    bound = "hbm"
    if dom_fp64 is not None and dom_fp64 > kern[dom]["hbm_frac"]:
        bound = "fp64"
    elif kern[dom].get("limiter"):
        bound = kern[dom]["limiter"]["unit"]
    roofline = {"bound": bound, "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None if traffic is None else traffic * 1e9,
                "traffic_source": tj.get("source"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": ms[dom],
                "fp64": kern[dom].get("fp64"), "limiter": kern[dom].get("limiter"),
                "kernels": kern,
                "floor_ms": {"hbm": hbm_floor_ms, "fp64": fp64_floor_ms,
                             "max": max(hbm_floor_ms, fp64_floor_ms or 0.0),
                             "note": "compulsory bytes of the step at the measured HBM peak; FP64 lane operations of the step "
                                     "(ncu, profiles/) at the measured DFMA peak (profiles/r2_dmma_ubench.txt)"},
                "path": {"algorithmic_bytes_per_step": path_alg, "achieved": path_alg / (ms_per_step * 1e-3) / 1e9,
                         "frac": path_alg / (ms_per_step * 1e-3) / 1e9 / peak,
                         "note": "compulsory bytes of update+formUnbalance+formTangent on this rank (state in/out, A and B "
                                 "out) over the whole step; the element records' round trip through HBM is overhead here"}}

    # ---- the linear solve, timed separately and out of path (north_star: left to the reference's SOE solver) ----
    solve = None
    if rank == 0 and a.solve_n > 0 and a.workload == "brick" and not a.no_cpu_baseline:
        try:
            import scipy.sparse as sp
            import scipy.sparse.linalg as spl
            from modelspec import J2_STEEL, brick_block
            sspec = brick_block(a.solve_n, a.solve_n, a.solve_n, mat=J2_STEEL)
            S = xb.DeviceModel.from_spec(sspec, xb.NUMBERER_RCM, xb.SOE_SPARSE_GEN_ROW).to_device(local)
            S.apply_load(1.0)
            As, Bs = S.form_tangent(), S.form_unbalance()
            sptr, sidx = S.pattern()
            M = sp.csr_matrix((As, sidx, sptr), shape=(S.neq, S.neq)).tocsc()
            t0s = time.perf_counter(); x = spl.spsolve(M, Bs); dts = time.perf_counter() - t0s
            solve = {"ms": dts * 1e3, "equations": int(S.neq), "nnz": int(S.nnz), "elements": int(sspec.ne),
                     "residual": float(np.abs(M @ x - Bs).max() / max(np.abs(Bs).max(), 1e-300)),
                     "solver": "scipy.sparse.linalg.spsolve (SuperLU, host, one core)",
                     "note": f"out of path and not optimised: the solve of a {a.solve_n}^3 block assembled by the device path (RCM numbering), "
                             "timed separately as north_star asks; the step numbers above contain no solve"}
            del S, M
        except Exception as ex:   # the solve is a side measurement: never let it take the bench line down
            solve = {"error": repr(ex)}

    # ---- end to end through the C-ABI with HOST buffers (pinned), copies inside the timed region ----
    e2e_steps = max(1, min(a.steps, a.e2e_steps)) if a.e2e_steps > 0 else 0
    # two trial fields, alternated, so that every e2e step is a genuine state determination (a force-based
    # beam returns at once when the displacement increment is zero)
    u_pin = torch.empty(u.size, dtype=torch.float64, pin_memory=True); u_pin.numpy()[:] = u.ravel()
    u_pin2 = torch.empty(u.size, dtype=torch.float64, pin_memory=True); u_pin2.numpy()[:] = 1.01 * u.ravel()
    A_pin = torch.empty(D.nnz, dtype=torch.float64, pin_memory=True)
    B_pin = torch.empty(max(D.nrows, 1), dtype=torch.float64, pin_memory=True)
    uns, An, Bn = (u_pin.numpy(), u_pin2.numpy()), A_pin.numpy(), B_pin.numpy()[:D.nrows]

    def e2e_step(i):
        D.set_trial_disp(uns[i % 2]); D.update(); D.form_unbalance(out=Bn); D.form_tangent(out=An)

    if e2e_steps:
        e2e_step(1)
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record(stream)
    for i in range(e2e_steps):
        e2e_step(i)
    e2.record(stream)
    barrier()
    e2e_ms = s2.elapsed_time(e2) / max(e2e_steps, 1)
    h2d, d2h = float(u.size * 8), float((D.nnz + D.nrows) * 8)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_ms = float(t[0])
        t = torch.tensor([h2d, d2h], device="cuda", dtype=torch.float64); dist.all_reduce(t); h2d, d2h = float(t[0]), float(t[1])
    # partition-invariant sums of the assembled system (every rank over its OWNED rows, summed over the ranks): the
    # same numbers whatever N is, up to the order of the final additions
    scale = None
    if e2e_steps:
        geq = D.row_eqns().astype(np.float64)
        part = np.array([An.sum(), np.abs(An).sum(), Bn.sum(), float(geq @ Bn), np.abs(Bn).sum()])
        if world > 1:
            t = torch.tensor(part, device="cuda", dtype=torch.float64); dist.all_reduce(t); part = t.cpu().numpy()
        inv = {"sum_A": float(part[0]), "sum_abs_A": float(part[1]), "sum_B": float(part[2]), "sum_eq_times_B": float(part[3]),
               "sum_abs_B": float(part[4])}
        key = f"{a.workload}:{n}"
        ref = load_invariants().get(key)
        if ref is None:
            scale = {"check": "no N=1 record for this size (tests/golden/bench_invariants.json)", "invariants": inv}
        else:
            tolA, tolB = 1e-12 * ref["sum_abs_A"], 1e-12 * ref["sum_abs_B"] * max(1.0, float(D.neq))
            bad = [k for k, tol in (("sum_A", tolA), ("sum_abs_A", tolA), ("sum_B", tolB), ("sum_abs_B", tolB), ("sum_eq_times_B", tolB))
                   if abs(inv[k] - ref[k]) > tol]
            scale = {"check": "ok" if not bad else "MISMATCH " + ",".join(bad), "invariants": inv, "n1_record": ref,
                     "tolerance": "1e-12 of sum|A| (A), 1e-12 of neq * sum|B| (B)"}
            if bad:
                print("bench.py: scale check FAILED: " + json.dumps(scale), file=sys.stderr)
    e2e = None if not e2e_steps else {"value": ngp_global / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps,
           "note": "xb_set_trial_disp(host u) + xb_update + xb_form_unbalance(host B) + xb_form_tangent(host A), pinned "
                   "host buffers; every rank moves its own nodes' u in and its owned rows of A, B out (bytes summed over ranks)"}

    # load balance of the partition: elements, owned rows and interface rows received per rank
    balance = None
    if world > 1:
        mine = torch.tensor([float(D.ne), float(D.nrows), float(sum(int(c[5]) for _, c in D.peers()))], device="cuda", dtype=torch.float64)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        arr = np.array([v.cpu().numpy() for v in allv])
        balance = {"elements_per_rank": arr[:, 0].astype(int).tolist(), "rows_per_rank": arr[:, 1].astype(int).tolist(),
                   "interface_chunks_in_per_rank": arr[:, 2].astype(int).tolist(),
                   "elements_max_over_mean": float(arr[:, 0].max() / arr[:, 0].mean()),
                   "interface_chunks_max_over_mean": float(arr[:, 2].max() / max(arr[:, 2].mean(), 1.0))}
    cb = None
    if not a.no_cpu_baseline and world == 1 and a.workload == "brick":
        cb, _ = cpu_arm(a.cpu_steps, 1, 1, a.cpu_sample)

    # ---- the other BASELINE configs on the same box, in the same line (N = 1 only): configs[1] 1000 x 1000 quads and
    # configs[3] the 195 k-element RC space frame under Newmark, each as a short run of this script ----
    secondary = None
    if a.secondary and world == 1 and a.workload == "brick" and not a.no_cpu_baseline:
        secondary = {}
        torch.cuda.synchronize()
        for wl in ("quad", "frame3d"):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", wl, "--steps", "5", "--warmup", "3",
                                    "--no-cpu-baseline", "--e2e-steps", "2"], capture_output=True, text=True, timeout=600)
                d = json.loads(r.stdout.strip().splitlines()[-1])
                secondary[wl] = {"workload": d["config"]["workload"], "elements": d["config"]["elements"], "gauss_points": d["config"]["gauss_points"],
                                 "ms_per_step": d["ms_per_step"], "value": d["value"], "unit": d["unit"],
                                 "kernel_ms": {k: v for k, v in d["kernel_ms"].items() if v > 0.005},
                                 "e2e_ms_per_step": d["e2e"]["ms_per_step"] if d.get("e2e") else None,
                                 "dominant": {"kernel": d["roofline"]["kernel"], "hbm_frac": d["roofline"]["frac"]}}
            except Exception as ex:
                secondary[wl] = {"error": repr(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n, a.workload), "elements": int(ne_global), "gauss_points": int(ngp_global),
                           "equations": int(D.neq), "rank0": {"elements": int(D.ne), "rows": int(D.nrows), "nnz": int(D.nnz),
                                                              "peers": [[int(r), int(c[0]), int(c[1])] for r, c in D.peers()]},
                           "partition": "none" if world == 1 else (
                               (f"METIS k-way ({world} parts, METIS_PartGraphKway of the reference's OTHER/METIS on the element graph, "
                                "computed on rank 0)" if a.partition == "metis" else f"recursive coordinate bisection, {world} parts")
                               + ", interface rows exchanged with NCCL send/recv"),
                           "balance": balance,
                           "numberer": "Plain", "soe": "SparseGenCol (CSC)",
                           "l2": "inputs larger than L2 (per GPU at N=1: state 7 GB, element matrices 19 GB, A 8 GB vs 126 MB)",
                           "setup_s": {"mesh": t_mesh, "host_setup": t_setup, "upload": t_upload}},
                "kernel_ms": ms_max, "kernel_ms_rank0": ms,
                "kernel_ms_note": "separate instrumented pass, one kernel after the other on one stream; in the timed region "
                                  "xb_form_tangent runs range by range on two streams (formTangent_call_ms), which hides the "
                                  "kernels' tails behind one another",
                "formTangent_ms": ms_max["element_tangent"] + ms_max["exchange_A"] + ms_max["assemble_A"],
                "formTangent_call_ms": ft_call_ms,
                "formUnbalance_ms": ms_max["element_resid"] + ms_max["exchange_B"] + ms_max["assemble_B"],
                "update_ms": ms_max["update"],
                "gpu_launches": int(launches_all), "clocks": clk, "e2e": e2e, "roofline": roofline, "cpu_baseline": cb,
                "scale_check": None if scale is None else scale["check"], "scale": scale, "solve_ms": solve, "secondary": secondary}
        json_out.write(json.dumps(line) + "\n"); json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=None, help="size parameter of the workload; default: brick 160 (4.096M elements), "
                                                        "quad 1000, frame 200, frame3d 20")
    ap.add_argument("--partition", default="rcb", choices=["rcb", "metis"],
                    help="N > 1: the library's recursive coordinate bisection, or METIS k-way as DomainPartitioner uses it")
    ap.add_argument("--workload", default="brick", choices=["brick", "quad", "frame", "frame3d", "soil"],
                    help="brick = the headline workload; quad / frame = secondary lines (profiles/), n = cells per side")
    ap.add_argument("--e2e-steps", type=int, default=3, help="0 skips the end-to-end leg (profiling runs only)")
    ap.add_argument("--cpu-sample", type=int, default=22, help="CPU arm sample: elements per side")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="name=value for xb_set_option (kernel tuning experiments)")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the short quad / frame3d runs whose numbers ride along in the brick line at N = 1")
    ap.add_argument("--selftest", action="store_true", help="torchrun, N >= 2: bitwise check of the NCCL-exchanged rows against "
                                                            "the single-GPU rows (prints {\"selftest\": \"ok\"})")
    ap.add_argument("--solve-n", type=int, default=16, help="size of the block whose linear solve is timed (out of path; 0 = skip)")
    a = ap.parse_args()
    if a.n is None:
        a.n = {"brick": 160, "quad": 1000, "frame": 200, "frame3d": 20, "soil": 100}[a.workload]
    if a.selftest:
        selftest_main(a)
    elif a.impl == "reference":
        reference_main(a)
    else:
        ours_main(a)


if __name__ == "__main__":
    main()
