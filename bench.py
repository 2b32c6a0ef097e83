#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (ICP NN queries/s & aligned hypotheses/s) on N B200s.

One "step" = one full multi-start ICP job over the workload's hypothesis batch (every rank runs the
same-sized shard: weak scaling, no data-path collective; the final all-gather of result records and the
arg-min are inside the timed region).  Default workload = BASELINE.json configs[2] (C3): 4096 initial
poses per GPU, 5k-point source vs 1M-point model, KNN-trimmed ICP, 30 iterations, grid NN.  configs[1]
(C2: one weighted alignment, 10k vs 500k, 100 iterations, brute-force NN) has a single hypothesis and
cannot shard; it is measured beside it at N=1 and reported under "c2" (with the brute-force kernel's
FP32-FMA roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c2g|c5|small]

--workload c5 = one GPU's share of configs[4] (2048 poses x 65 536 source points vs a 16M-point model that is not
L2-resident; ~8 s per step, so run it with --steps 2 --no-c2 --no-match --no-cpu).

--impl reference times the CPU oracle restatement (the reference is MATLAB; MATLAB/Octave are probed
and reported, neither exists in this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FFMA_PEAK_TFLOPS_MEASURED = 65.7      # tools/fma_peak.cu on this pool's B200 (profiles/r01_fma_peak.txt)
HBM_FALLBACK_GBS = 6650.0             # /opt/skills/guides/B200_PROFILING.md fallback

WORKLOADS = {
    # name: nm, ns, hyps per GPU, iters, mode, nn, pose grid (n_rot, (nx,ny,nz)), max_deg, sigma, seed
    "c3": dict(nm=1_000_000, ns=5000, hyp=4096, iters=30, mode="knn", nn="grid", rot=16, trans=(8, 8, 4), max_deg=20.0,
               sigma=0.3, seed=1003, desc="C3 multi-start ICP: 4096 poses/GPU x 5k src vs 1M model, KNN-trimmed, 30 it, grid NN"),
    "c2": dict(nm=500_000, ns=10_000, hyp=1, iters=100, mode="weighted", nn="brute", rot=1, trans=(1, 1, 1), max_deg=0.0,
               sigma=0.3, seed=1002, desc="C2 AlignPoints_weighted-style single alignment: 10k weighted src vs 500k model, 100 it, brute NN"),
    "c2g": dict(nm=500_000, ns=10_000, hyp=1, iters=100, mode="weighted", nn="grid", rot=1, trans=(1, 1, 1), max_deg=0.0,
                sigma=0.3, seed=1002, desc="C2 with the grid NN path (same inputs and results as C2, what a user would run)"),
    # one GPU's share of BASELINE.json configs[4] (16k hypotheses over 8 GPUs): the model (512 MB of points + grid) is NOT
    # L2-resident, the row scan runs as the warp-per-query kernel (dense model).  Optional: `--workload c5`, ~7 s per step.
    "c5": dict(nm=16_000_000, ns=65536, hyp=2048, iters=20, mode="knn", nn="grid", rot=8, trans=(8, 8, 4), max_deg=10.0,
               sigma=0.3, seed=1005, src_stride=16,
               desc="C5 large upsampled model: 2048 poses/GPU x 65 536 src vs 16M model, KNN-trimmed, 20 it, grid NN"),
    "small": dict(nm=100_000, ns=2000, hyp=256, iters=10, mode="knn", nn="grid", rot=4, trans=(4, 4, 4), max_deg=10.0,
                  sigma=0.3, seed=7, desc="small multi-start ICP (debug)"),
}


def make_inputs(w, rank, world=1):
    """Synthetic inputs of workload w for one rank.  Multi-GPU (weak scaling): ONE pose grid of hyp * world poses --
    the same rotations, the z lattice refined `world` times over the same extent -- dealt to the ranks round-robin,
    so that every GPU gets a statistically identical share."""
    from pcreg_b200 import synth
    model = synth.make_model(w["nm"], w["seed"])
    src, T_gt, c = synth.make_source(model[::w.get("src_stride", 1)], w["ns"], w["sigma"], w["seed"])
    if w["hyp"] == 1:
        T0 = synth.perturb_pose(T_gt, c, synth.rot_axis_angle([0.3, -0.5, 0.8], np.deg2rad(5.0)), np.array([1.2, -1.0, 1.2]))[None]
    else:
        nx, ny, nz = w["trans"]
        T0 = synth.pose_grid(T_gt, c, w["rot"], (nx, ny, nz * world), w["max_deg"], (2.0, 2.0, 2.0 / world), w["seed"])
        T0 = np.ascontiguousarray(T0[rank::world][: w["hyp"]])
    g = synth.rng(w["seed"] + 5)
    w_src = g.uniform(0.5, 1.0, w["ns"]) if w["mode"] == "weighted" else None
    return model, src, T0, w_src, T_gt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        self.proc = subprocess.Popen([exe, "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples in the timed region"], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, STREAM-style copy)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get(kernel)
        except Exception:
            return None
    return None


def probe_matlab():
    found = [x for x in ("matlab", "octave", "octave-cli") if shutil.which(x)]
    return found


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle restatement): bounded sample of the same workload
# ----------------------------------------------------------------------------------------------
def cpu_sample(w, budget_s=12.0, max_hyp=None):
    import oracle
    model, src, T0, w_src, _ = make_inputs(w, 0)
    omode = dict(plain=oracle.ICP_PLAIN, knn=oracle.ICP_KNN, weighted=oracle.ICP_WEIGHTED)[w["mode"]]
    t_build0 = time.perf_counter()
    nn = oracle.nn.KDTreeNN(np.asarray(model, dtype=np.float64))
    t_build = time.perf_counter() - t_build0
    n = 0
    t0 = time.perf_counter()
    limit = T0.shape[0] if max_hyp is None else min(max_hyp, T0.shape[0])
    while n < limit:
        oracle.icp_single(model, src, T0[n], mode=omode, iters=w["iters"], k_frac=0.85, R_w=3.5, w_src=w_src, nn=nn)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    q = n * w["ns"] * (w["iters"] + 1)
    return dict(queries_per_s=q / dt, hyp_per_s=n / dt, n_hyp=n, seconds=dt, kdtree_build_s=t_build)


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        cpu_sample(w, budget_s=1.0, max_hyp=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_sample(w, budget_s=max(2.0, 60.0 / max(1, args.steps))))
    dt = time.perf_counter() - t0
    v = float(np.mean([x["queries_per_s"] for x in vals]))
    sample = "%d hypotheses/step of the %s workload (oracle restatement: numpy FP64 + scipy cKDTree workers=-1; %s)" % (
        vals[-1]["n_hyp"], args.workload,
        "MATLAB/Octave not installed" if not probe_matlab() else "found " + ",".join(probe_matlab()) + " but the composed ICP has no .m file")
    line = dict(impl="reference", metric="ICP NN queries/s", value=v, unit="queries/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / max(1, args.steps), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=w["desc"], l2="n/a (CPU)"),
                hyp_per_s=float(np.mean([x["hyp_per_s"] for x in vals])),
                cpu_baseline=dict(value=v, unit="queries/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def gpu_workload(P, torch, w, rank, steps, warmup, dist=None, world=1, e2e_steps=None):
    """Returns a dict of measurements for one workload on this rank (collectives included when world > 1)."""
    from pcreg_b200 import sharded, torch_ops
    dev = torch.device("cuda", torch.cuda.current_device())
    # PCREG_BENCH_SHARD="r/n": measure on ONE GPU the share that rank r of an n-GPU run would get (work-balance check)
    shard = os.environ.get("PCREG_BENCH_SHARD")
    in_rank, in_world = (int(shard.split("/")[0]), int(shard.split("/")[1])) if shard else (rank, world)
    model_h, src, T0, w_src, T_gt = make_inputs(w, in_rank, in_world)
    m = P.Model(model_h, grid=(w["nn"] == "grid"), cells_per_point=float(os.environ.get("PCREG_GRID_CPP", "0")))
    mode = dict(plain=P.ICP_PLAIN, knn=P.ICP_KNN, weighted=P.ICP_WEIGHTED)[w["mode"]]
    nn = P.NN_GRID if w["nn"] == "grid" else P.NN_BRUTE
    opts = P.icp_opts(mode=mode, iters=w["iters"], k_frac=0.85, R_w=3.5, nn=nn)
    H, ns = T0.shape[0], src.shape[0]
    # device-resident inputs
    src_cm = torch_ops.src_to_abi_t(torch.from_numpy(src).to(dev))
    T0_abi = torch_ops.T_to_abi_t(torch.from_numpy(T0).to(dev))
    w_t = torch.from_numpy(w_src).to(dev) if w_src is not None else None
    out = torch_ops.IcpDeviceBuffers(H, ns, w["iters"], dev)
    per = H

    def step_device():
        torch_ops.icp_batch_device(m, src_cm, w_t, T0_abi, opts, out)
        if world > 1:
            rec = torch.cat([out.rmse[:, None], out.T, out.n_used[:, None].double(), out.status[:, None].double()], dim=1)
            allrec, best = sharded.gather_and_pick(rec, H * world, per)
            return best
        return int(out.best.item())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step_device()
    sync_all()
    launches0 = P.launch_count()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    sync_all()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = P.launch_count() - launches0
    # one extra, UNTIMED step in profiling mode: per-kernel CUDA-event times and exact device-side work counters
    # (the counters add atomics to the kernels, so the timed steps above run without them)
    P.set_profiling(True)
    step_device()
    sync_all()
    prof = P.last_profile()
    P.set_profiling(False)
    ms_ranks = [ms / steps]
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        allms = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allms, tt)
        ms_ranks = [float(x) / steps for x in allms.cpu()]
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    rmse_best = float(out.rmse[int(out.best.item())].item())

    # ---- e2e: the public host-buffer API (pinned host inputs, H2D + D2H inside the timed region) ----
    e2e_steps = e2e_steps if e2e_steps is not None else steps
    import ctypes as C
    from pcreg_b200 import _lib as L
    h_src = torch.from_numpy(np.asfortranarray(src).T.copy()).pin_memory()            # [3, ns] = column-major ns x 3
    h_T0 = torch.from_numpy(np.ascontiguousarray(np.swapaxes(T0, 1, 2))).pin_memory()
    h_w = torch.from_numpy(w_src).pin_memory() if w_src is not None else None
    h_T = torch.empty((H, 16), dtype=torch.float64).pin_memory()
    h_rmse = torch.empty(H, dtype=torch.float64).pin_memory()
    h_nu = torch.empty(H, dtype=torch.int32).pin_memory()
    h_st = torch.empty(H, dtype=torch.int32).pin_memory()
    best = C.c_int64()
    vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    cast = lambda t, ty: C.cast(C.c_void_p(t.data_ptr()), ty) if t is not None else None

    def step_host():
        L.check(L.lib().pcreg_icp_batch(m.handle, vp(h_src), 1, ns, ns, cast(h_w, L.c_f64p), cast(h_T0, L.c_f64p), H, C.byref(opts),
                                        cast(h_T, L.c_f64p), cast(h_rmse, L.c_f64p), cast(h_nu, L.c_i32p), cast(h_st, L.c_i32p),
                                        None, None, C.byref(best)), "pcreg_icp_batch")
        if world > 1:
            rec = torch.cat([h_rmse[:, None], h_T, h_nu[:, None].double(), h_st[:, None].double()], dim=1).to(dev)
            sharded.gather_and_pick(rec, H * world, per)

    step_host()
    sync_all()
    e0.record()
    for _ in range(e2e_steps):
        step_host()
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    h2d = h_src.numel() * 8 + h_T0.numel() * 8 + (h_w.numel() * 8 if h_w is not None else 0)
    d2h = h_T.numel() * 8 + h_rmse.numel() * 8 + h_nu.numel() * 4 + h_st.numel() * 4 + 8
    ok_e2e = bool(np.allclose(h_rmse.numpy(), out.rmse.cpu().numpy(), rtol=0, atol=0, equal_nan=True))
    grid = m.grid_info() if w["nn"] == "grid" else None
    if grid is not None:
        grid["voxel_map"] = m.voxel_info()
    m.destroy()
    q_per_step = H * ns * (w["iters"] + 1)
    return dict(ms_per_step=ms / steps, ms_per_step_e2e=ms_e2e / e2e_steps, q_per_step=q_per_step, H=H, ns=ns, launches=launches,
                prof=prof, clocks=clocks, ms_ranks=ms_ranks, h2d=h2d, d2h=d2h, rmse_best=rmse_best, e2e_equals_device=ok_e2e, grid=grid,
                nm=model_h.shape[0])


def roofline_for(w, r):
    """Roofline of the dominant kernel of the workload, from the profiled step (CUDA events around every launch
    inside the library + exact device-side work counters).  Algorithmic bytes as defined in DESIGN.md section 3."""
    p = r["prof"]
    if w["nn"] == "grid" and p.get("voxel_map"):
        # Voronoi voxel map: per query 24-byte source point + 8-byte voxel header + 12-byte result, 16 bytes per list entry
        # scanned, 32 bytes per point gathered for the FP64 decision; the rest of the queries is walked (counters as below)
        peak, how = hbm_peak()
        kern = {
            "k_nn_vox": dict(ms=p["list_ms"], launches=p["list_launches"],
                             bytes=8.0 * p["nn_queries"] + 36.0 * p["certified_queries"] + 16.0 * p["list_entries_read"]
                                   + 32.0 * p["list_points_gathered"]),
            "k_nn_grid_walk": dict(ms=p["walk_ms"], launches=p["walk_launches"],
                                   bytes=36.0 * p["walked_queries"] + 8.0 * p["walk_leaves"] + 32.0 * p["walk_points"]
                                         + 1.0 * max(0.0, p["grid_nodes_popped"] - p["walk_leaves"])),
            "k_icp_update": dict(ms=p["update_ms"], launches=p["update_launches"], bytes=76.0 * p["correspondences"]),
        }
        for k, v in kern.items():
            v["gbs"] = v["bytes"] / max(v["ms"], 1e-9) / 1e6
            v["frac"] = v["gbs"] / peak
            v["traffic"] = ncu_traffic(k)
        top = max(kern, key=lambda k: kern[k]["ms"])
        t = kern[top]
        launches = max(1.0, t["launches"])
        return dict(bound="hbm", kernel=top, achieved=t["gbs"], peak=peak, unit="GB/s", frac=t["frac"], traffic=t["traffic"],
                    peak_source=how, bytes_per_launch=t["bytes"] / launches, avg_launch_ms=t["ms"] / launches,
                    note="algorithmic bytes from exact device-side counters (DESIGN.md 3.2)",
                    kernels={k: dict(ms=v["ms"], launches=v["launches"], algorithmic_bytes=v["bytes"], gbs=v["gbs"], frac=v["frac"],
                                     traffic=v["traffic"]) for k, v in kern.items()},
                    list_answered_fraction=p["certified_queries"] / max(1.0, p["nn_queries"]),
                    entries_per_query=p["list_entries_read"] / max(1.0, p["certified_queries"]),
                    fp64_points_per_query=p["list_points_gathered"] / max(1.0, p["certified_queries"]))
    if w["nn"] == "grid":
        peak, how = hbm_peak()
        nq_pass = r["H"] * r["ns"]
        kern = {
            # every query of the pass reads its 8-byte list descriptor; a scanned query also its 16-byte header, its
            # 24-byte source point and writes 12 bytes; 4 bytes per list entry read, 32 bytes per model point gathered
            "k_nn_list": dict(ms=p["list_ms"], launches=p["list_launches"],
                              bytes=8.0 * nq_pass * p["list_launches"] + 52.0 * p["certified_queries"] + 4.0 * p["list_entries_read"]
                                    + 32.0 * p["list_points_gathered"], queries=p["certified_queries"]),
            # 24-byte source point + 12-byte result per query, 8 bytes (start, end) per visited cell row, 32 bytes per point
            "k_nn_grid_direct": dict(ms=p["rowscan_ms"], launches=p["rowscan_launches"],
                                     bytes=36.0 * p["rowscan_queries"] + 8.0 * p["rowscan_rows"] + 32.0 * p["rowscan_points"],
                                     queries=p["rowscan_queries"] - p["walked_queries"] + nq_pass),
            # same per query / leaf / point, plus 1 mask byte per expanded pyramid node
            "k_nn_grid_walk": dict(ms=p["walk_ms"], launches=p["walk_launches"],
                                   bytes=36.0 * p["walked_queries"] + 8.0 * p["walk_leaves"] + 32.0 * p["walk_points"]
                                         + 1.0 * max(0.0, p["grid_nodes_popped"] - p["walk_leaves"]), queries=p["walked_queries"]),
        }
        resident = r["nm"] * 64 <= 126e6            # points (2 x 32 B) + grid fit the 126 MB L2: the committed C3 captures apply
        for k, v in kern.items():
            v["gbs"] = v["bytes"] / max(v["ms"], 1e-9) / 1e6
            v["frac"] = v["gbs"] / peak
            v["traffic"] = ncu_traffic(k) if resident else None
        top = max(kern, key=lambda k: kern[k]["ms"])
        t = kern[top]
        launches = max(1.0, t["launches"])
        return dict(bound="hbm", kernel=top, achieved=t["gbs"], peak=peak, unit="GB/s", frac=t["frac"], traffic=t["traffic"],
                    peak_source=how, bytes_per_launch=t["bytes"] / launches, avg_launch_ms=t["ms"] / launches,
                    note=("algorithmic bytes from exact device-side counters (DESIGN.md 3.2); the 1M-point model and its grid are "
                          "L2-resident, so the gathers are served by L2 and DRAM traffic (traffic, from ncu) is far below it: the "
                          "kernels are bound by gather latency / L1 wavefronts, not by HBM") if resident else
                         ("algorithmic bytes from exact device-side counters (DESIGN.md 3.2); the model is not L2-resident, every visited "
                          "point comes from HBM (profiles/r01_ncu_full_c5_rows.txt: DRAM bytes = algorithmic bytes for the row scan); "
                          "k_nn_grid_direct stands for the row-scan kernel that ran (k_nn_grid_rows on dense models)"),
                    kernels={k: dict(ms=v["ms"], launches=v["launches"], algorithmic_bytes=v["bytes"], gbs=v["gbs"], frac=v["frac"],
                                     traffic=v["traffic"]) for k, v in kern.items()},
                    list_answered_fraction=p["certified_queries"] / max(1.0, p["nn_queries"]),
                    points_visited_per_query=(p["grid_points_visited"] + p["list_points_gathered"]) / p["nn_queries"])
    pairs, launches = p["brute_pairs"], max(1.0, p["nn_launches"])
    achieved = 6.0 * pairs / (p["nn_ms"] * 1e-3) / 1e12
    return dict(bound="fp32_fma", kernel="k_nn_brute", achieved=achieved, peak=FFMA_PEAK_TFLOPS_MEASURED, unit="TFLOP/s",
                frac=achieved / FFMA_PEAK_TFLOPS_MEASURED, traffic=ncu_traffic("k_nn_brute"),
                peak_source="measured FFMA microbenchmark tools/fma_peak.cu (theoretical 74.4 TFLOP/s at 1965 MHz)",
                flops_per_launch=6.0 * pairs / launches, avg_launch_ms=p["nn_ms"] / launches,
                note="6 FLOP per (query, model point) pair; includes the bound pass and the FP64 slow path in the time")


FP64_INST_PEAK_MEASURED = 16.45e12    # DFMA issue rate, tools/fma_peak.cu (32.9 TFLOP/s; profiles/r01_fma_peak.txt)


def match_workload(P, torch, n1=3000, n2=20000, dim=980, reps=3):
    """SURVEY.md section 8 row f4 beside the headline: getMatches (descriptor weighting + exhaustive matchFeatures) of n1 surface
    against n2 model descriptors through the host-buffer API.  The score kernel is FP64-pipe bound: 2 FP64 instructions
    per (pair, dimension) term for SAD."""
    g = np.random.default_rng(1)
    dS = g.poisson(g.gamma(0.6, 4.0, (n1, dim))).astype(np.float64)
    dM = g.poisson(g.gamma(0.6, 4.0, (n2, dim))).astype(np.float64)
    dM[::7][: n1 // 2] = dS[: n1 // 2]                                             # some exact partners, so matches exist
    par = dict(UNNORMALIZE=True, norm_factor=2, CHANGE_METRIC=True, metric_factor=0.6, MatchThreshold=10, MaxRatio=0.99,
               Metric="SAD", Unique=True)                                          # completeExperiment.m:112-122
    dS, dM = np.asfortranarray(dS), np.asfortranarray(dM)
    P.getMatches(dS[:64], dM[:64], par)
    launches0 = P.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        matches = P.getMatches(dS, dM, par)
    wall_ms = (time.perf_counter() - t0) * 1e3 / reps
    launches = (P.launch_count() - launches0) // reps
    P.set_profiling(True)
    P.getMatches(dS, dM, par)
    pr = P.last_profile()
    P.set_profiling(False)
    inst_s = 2.0 * pr["match_terms"] / (pr["match_score_ms"] * 1e-3)
    return dict(workload="getMatches: %d surface x %d model descriptors, 981 dimensions, SAD, exhaustive matchFeatures" % (n1, n2),
                value=n1 * n2 / (wall_ms * 1e-3), unit="descriptor pairs/s (host buffers, H2D inside)", ms_per_call=wall_ms,
                h2d_bytes_per_call=int(dS.nbytes + dM.nbytes), matches=int(matches.shape[0]), gpu_launches=int(launches),
                roofline=dict(bound="fp64_pipe", kernel="k_match_scores", achieved=inst_s / 1e12, peak=FP64_INST_PEAK_MEASURED / 1e12,
                              unit="T FP64 inst/s", frac=inst_s / FP64_INST_PEAK_MEASURED, avg_launch_ms=pr["match_score_ms"],
                              terms_per_launch=pr["match_terms"],
                              note="2 FP64 instructions per (pair, dimension) term (t = a - b; acc += |t|); peak = measured DFMA issue rate"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-c2", action="store_true", help="skip the secondary C2 (brute-force) measurement at N=1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-match", action="store_true", help="skip the getMatches (row f4) side measurement at N=1")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import pcreg_b200 as P
    P.init(local_rank)

    r = gpu_workload(P, torch, w, rank, args.steps, max(3, args.warmup) if args.warmup >= 0 else 3, dist, world)
    sec = r["ms_per_step"] * 1e-3
    value = r["q_per_step"] * world / sec
    line = dict(metric="ICP NN queries/s", value=value, unit="queries/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=w["desc"], hypotheses_per_gpu=r["H"], source_points=r["ns"], model_points=r["nm"], iters=w["iters"],
                            parallelism="hypotheses sharded, model replicated, final all-gather + arg-min" if world > 1 else "single GPU",
                            l2=("no flush: per-step correspondence scratch (%.0f MB) exceeds the 126 MB L2; " % (r["H"] * r["ns"] * 24 / 1e6))
                               + ("the 1M-point model is L2-resident by design" if r["nm"] * 64 <= 126e6 else
                                  "the model (%.0f MB of points + grid) exceeds it as well" % (r["nm"] * 64 / 1e6)), grid=r["grid"]),
                hyp_per_s=r["H"] * world / sec,
                e2e=dict(value=r["q_per_step"] * world / (r["ms_per_step_e2e"] * 1e-3), unit="queries/s", h2d_bytes_per_step=r["h2d"],
                         d2h_bytes_per_step=r["d2h"], ms_per_step=r["ms_per_step_e2e"], hyp_per_s=r["H"] * world / (r["ms_per_step_e2e"] * 1e-3),
                         result_equals_device_path=r["e2e_equals_device"]),
                gpu_launches=r["launches"], clocks=r["clocks"], ms_per_step_by_rank=r["ms_ranks"], roofline=roofline_for(w, r),
                kernel_time_share=dict(nn_ms=r["prof"]["nn_ms"], update_ms=r["prof"]["update_ms"], list_ms=r["prof"]["list_ms"],
                                       rowscan_ms=r["prof"]["rowscan_ms"], walk_ms=r["prof"]["walk_ms"], step_ms=r["ms_per_step"],
                                       note="from one extra profiled step (event records + counters), not from the timed steps"),
                best_rmse=r["rmse_best"])
    if rank == 0 and world == 1 and not args.no_c2 and args.workload == "c3":
        w2 = WORKLOADS["c2"]
        r2 = gpu_workload(P, torch, w2, 0, max(2, args.steps), 3)
        s2 = r2["ms_per_step"] * 1e-3
        line["c2"] = dict(workload=w2["desc"], value=r2["q_per_step"] / s2, unit="queries/s", ms_per_step=r2["ms_per_step"],
                          hyp_per_s=1.0 / s2, e2e=dict(value=r2["q_per_step"] / (r2["ms_per_step_e2e"] * 1e-3), unit="queries/s",
                                                       ms_per_step=r2["ms_per_step_e2e"]),
                          roofline=roofline_for(w2, r2), gpu_launches=r2["launches"], best_rmse=r2["rmse_best"], clocks=r2["clocks"])
        w3 = WORKLOADS["c2g"]
        r3 = gpu_workload(P, torch, w3, 0, max(2, args.steps), 3)
        line["c2"]["grid_path"] = dict(workload=w3["desc"], value=r3["q_per_step"] / (r3["ms_per_step"] * 1e-3), unit="queries/s",
                                       ms_per_step=r3["ms_per_step"], e2e_ms_per_step=r3["ms_per_step_e2e"], best_rmse=r3["rmse_best"],
                                       same_result_as_brute=bool(r3["rmse_best"] == r2["rmse_best"]))
    if rank == 0 and world == 1 and not args.no_match and args.workload == "c3":
        line["get_matches"] = match_workload(P, torch)
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        c = cpu_sample(w, budget_s=15.0)
        line["cpu_baseline"] = dict(value=c["queries_per_s"], unit="queries/s", cores=cores, kind="port",
                                    hyp_per_s=c["hyp_per_s"],
                                    sample="%d of the %d hypotheses of the same workload, full %d iterations each (%.1f s; oracle restatement: "
                                           "numpy FP64 + scipy cKDTree workers=-1, tree build %.1f s excluded; MATLAB/Octave %s)"
                                           % (c["n_hyp"], r["H"], w["iters"], c["seconds"], c["kdtree_build_s"],
                                              "not installed" if not probe_matlab() else "present: " + ",".join(probe_matlab())))
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
