#!/usr/bin/env python3
"""bench.py — model-to-scene registrations/s (BASELINE.json metric) on N B200s, plus ICP iterations/s @ 1 M points.

A step = one pass of the hot path over one batch: every model of the database (BASELINE.json configs[1]: chair1, chair2,
chair4, desk1, desk2, desk3, sofa, Chair_025 x0.01) registered against mcloud.pcd — grid -> normals -> Harris -> FPFH
(both clouds) -> feature k-NN -> 50 000 prerejective RANSAC hypotheses -> ICP (PCL defaults) — every stage recomputed
in every registration (cached stages are dropped with rtr_cloud_reset before each one).

  value : registrations/s with the clouds already resident in HBM (rtr_register on device handles)
  e2e   : the same through the host-buffer entry point (rtr_register_host: pinned host clouds in, H2D + D2H inside)
  N > 1 : model-sharded, weak scaling — every rank registers its own shard of 8 models (database of 8 N models) against
          the replicated scene, then ONE all-gather of the 128-byte pose records (NCCL), inside the timed region.
  --impl reference : the CPU oracle (oracle/liboracle.so, all host threads) on the same workload.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from realtime_robot_b200.params import default_register_params  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402

MODELS = ["chair1", "chair2", "chair4", "desk1", "desk2", "desk3", "sofa", "Chair_025"]
SCENE = "mcloud"
METRIC = "model-to-scene registrations/s"
UNIT = "registrations/s"


def load_cloud(name):
    src = "desk1" if name == "desk2" else name          # desk2.pcd is byte-identical to desk1.pcd in the reference
    pts = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", src + ".pcd")))
    if name == "Chair_025":
        pts[:, :3] *= np.float32(0.01)                  # units x100; the scale model_point.h:106-111 intends
    return pts


def workload_config(params):
    return {"workload": "configs[1]: all repo models (chair1, chair2, chair4, desk1, desk2, desk3, sofa, Chair_025 x0.01) vs "
                        "mcloud.pcd, full pipeline per registration (normals r=.05, Harris r=.05 thr=.01 NMS+refine, FPFH r=.10, "
                        "k-NN k=5, prerejective RANSAC, ICP PCL defaults)",
            "registrations_per_step_per_gpu": len(MODELS),
            "ransac_hypotheses": int(params.ransac.max_iterations),
            "icp_iterations": int(params.icp.max_iterations),
            "parallelism": "model-sharded, 8 models per rank queued on 8 streams (rtr_register_begin / _end), one 128 B/record all-gather per step",
            "l2": "flushed between timed steps (256 MiB write); inputs are < 1 MB"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    # every host core this process may use (torchrun exports OMP_NUM_THREADS=1, which omp_get_max_threads() would obey)
    cores = orc.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    p = default_register_params()
    scene = load_cloud(SCENE)
    models = [(m, load_cloud(m)) for m in MODELS]
    t = time.time()
    for _, m in models:
        orc.register(m, scene, p)
    full = time.time() - t
    budget = 150.0
    sample = models
    if full * (args.steps + args.warmup - 1) > budget:
        # keep the run bounded: smallest-first subset that fits
        order = sorted(models, key=lambda kv: len(kv[1]))
        per_pt = full / sum(len(m) for _, m in models)
        sample, est = [], 0.0
        for kv in order:
            est += per_pt * len(kv[1]) * (args.steps + args.warmup)
            if sample and est > budget:
                break
            sample.append(kv)
    for _ in range(max(args.warmup - 1, 0)):
        for _, m in sample:
            orc.register(m, scene, p)
    t0 = time.time()
    for _ in range(args.steps):
        for _, m in sample:
            orc.register(m, scene, p)
    dt = time.time() - t0
    value = len(sample) * args.steps / dt
    desc = f"{len(sample)} of the step's {len(MODELS)} registrations ({', '.join(n for n, _ in sample)}) per step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "repo .pcd clouds (xyz copies under data/clouds)",
            "config": workload_config(p),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "PCL 1.8.0 cannot be built here (not vendored, not installed): the reference arm is the CPU oracle, a "
                    "restatement of the same algorithms, run with all host threads"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); out["sm_max_mhz"] = float(r[1])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[3 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-icp", action="store_true", help="skip the ICP @ 1M-point and PCD I/O sections")
    ap.add_argument("--host-threads", type=int, default=2, help="host threads per rank that queue the step's 8 registrations "
                    "(rtr_register_begin / _end); 8 = one thread per registration")
    ap.add_argument("--verbose", action="store_true", help="per-step event / wall times on stderr")
    ap.add_argument("--no-native", action="store_true", help="skip the reference-native descriptor path section")
    ap.add_argument("--no-scene", action="store_true", help="skip the 4M-point scene section (configs[3]: normals + FPFH)")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its version banner there)
    import torch
    import torch.distributed as td
    from realtime_robot_b200 import api, dist, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the registration path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    # one context (stream) per in-flight registration: the step's 8 registrations are issued concurrently from 8 host
    # threads (ctypes releases the GIL), so small kernels of different registrations share the 148 SMs
    from concurrent.futures import ThreadPoolExecutor
    ctxs = [api.Context(local_rank) for _ in MODELS]
    ctx = ctxs[0]
    pool = ThreadPoolExecutor(max_workers=len(MODELS))
    p = default_register_params()
    hbm_peak, bf16_peak, peak_src = peaks()
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        ncu_traffic = {}

    # host clouds in pinned memory (e2e leg) and resident device clouds (value leg)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    scene_t, scene_h = pinned(load_cloud(SCENE))
    models_h = []
    keep = [scene_t]
    for m in MODELS:
        t, h = pinned(load_cloud(m))
        keep.append(t)
        models_h.append(h)
    scenes_d = [api.Cloud(c, scene_h) for c in ctxs]
    scene_d = scenes_d[0]
    models_d = [api.Cloud(c, h) for c, h in zip(ctxs, models_h)]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def gather(recs):
        for i, r in enumerate(recs):
            r.model_id = rank * len(MODELS) + i
        if world == 1:
            return recs
        return dist.all_gather_records(recs, len(MODELS), device=dev)

    def reg_resident(i):
        models_d[i].reset(); scenes_d[i].reset()
        return api.register(models_d[i], scenes_d[i], p)

    def reg_e2e(i):
        return api.register_host(ctxs[i], models_h[i], scene_h, p)

    # issue order: the longest registration first (chair4), so that its chain starts at once and the short ones fill in
    order = sorted(range(len(MODELS)), key=lambda i: -models_d[i].n)

    # T host threads per rank, each queues its share of the 8 registrations on their streams (rtr_register_begin does not
    # synchronise) and then collects the records: T = 1 needs one core per rank, T = 8 is one synchronous call per thread
    T = max(1, min(args.host_threads, len(MODELS)))
    groups = [order[w::T] for w in range(T)]

    def run_group(w, begin):
        for i in groups[w]:
            begin(i)
        return [(i, api.register_end(ctxs[i])) for i in groups[w]]

    def run_step(begin):
        parts = [run_group(0, begin)] if T == 1 else list(pool.map(lambda w: run_group(w, begin), range(T)))
        recs = [None] * len(MODELS)
        for part in parts:
            for i, r in part:
                recs[i] = r
        return gather(recs)

    def begin_resident(i):
        models_d[i].reset(); scenes_d[i].reset()
        api.register_begin(models_d[i], scenes_d[i], p)

    def step_resident():
        return run_step(begin_resident)

    def step_e2e():
        return run_step(lambda i: api.register_host_begin(ctxs[i], models_h[i], scene_h, p))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        """Device time of `steps` steps: CUDA events on the library's own stream around the registrations of each step,
        plus torch events around the all-gather; L2 flushed (untimed) between steps."""
        total_ms, last = 0.0, None
        for _ in range(steps):
            flush_buf.fill_(1)
            torch.cuda.synchronize()
            ctx.record(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            last = step_fn()
            ctx.record(1)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            ev = ctx.elapsed_ms(0, 1)
            # the registrations are synchronous at their end (result D2H), and the all-gather runs on torch's stream after
            # them: the step's device time is bounded below by the event span and above by the synced wall span
            total_ms += max(ev, wall) if world > 1 else ev
            if args.verbose and rank == 0:
                print(f"[bench] step: events {ev:.3f} ms, host wall {wall:.3f} ms", file=sys.stderr, flush=True)
        return total_ms, last

    def log(msg):
        if rank == 0:
            print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)

    log("warm-up")
    for _ in range(args.warmup):
        t0 = time.perf_counter(); step_resident(); t1 = time.perf_counter(); step_e2e()
        log(f"  warm-up step: resident {1e3 * (t1 - t0):.1f} ms, e2e {1e3 * (time.perf_counter() - t1):.1f} ms")

    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = sum(c.launches for c in ctxs)
    wall0 = time.perf_counter()
    ms_res, recs = timed(step_resident, args.steps)
    barrier()
    wall_res = time.perf_counter() - wall0
    launches = sum(c.launches for c in ctxs) - l0
    ms_e2e, recs_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop() if sampler else None

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_res, ms_e2e = float(t[0]), float(t[1])
    n_reg = len(MODELS) * world * args.steps
    value = n_reg / (ms_res * 1e-3)
    e2e_value = n_reg / (ms_e2e * 1e-3)
    h2d = sum((len(h) + len(scene_h)) * 16 for h in models_h)
    d2h = 128 * len(MODELS)

    log(f"timed: resident {ms_res / args.steps:.2f} ms/step, e2e {ms_e2e / args.steps:.2f} ms/step")
    # ---- per-kernel device time of one more step (profiling marks; not part of the timed region)
    # (run one registration at a time here so that the intervals of different streams do not overlap)
    prof = {}
    per_model_ms = {}
    for i in range(len(MODELS)):
        ctxs[i].profile_begin()
        reg_resident(i)
        pi = ctxs[i].profile_end()
        per_model_ms[MODELS[i]] = round(sum(v[1] for v in pi.values()), 3)
        for k, v in pi.items():
            a = prof.get(k, (0, 0.0))
            prof[k] = (a[0] + v[0], a[1] + v[1])
    # one registration at a time, unprofiled: latency of a single registration as a caller sees it
    lat = {}
    for i in range(len(MODELS)):
        reg_resident(i)
        t0 = time.perf_counter(); reg_resident(i); lat[MODELS[i]] = round(1e3 * (time.perf_counter() - t0), 3)
    step_ms = sum(v[1] for v in prof.values())      # serialised device time of the step's operations
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])
    kernel_share = {k: {"launches": v[0], "ms": round(v[1], 4), "share": round(v[1] / step_ms, 4)} for k, v in top[:8]}

    # algorithmic bytes of the dominant kernel (DESIGN.md "Kernels and their rooflines")
    def neighbour_sums(r):
        tot = []
        for cm in models_d:
            cnt, _, _ = cm.radius_neighbors(r, counts_only=True)
            cs, _, _ = scene_d.radius_neighbors(r, counts_only=True)
            tot.append((int(cnt.sum()), cm.n, int(cs.sum()), scene_d.n))
        return tot
    dom = top[0][0]
    alg = None
    if dom == "ransac.eval":
        alg = sum(int(r.evaluated) * (cm.n * 32 + 64) for r, cm in zip(recs[:len(MODELS)] if world == 1 else recs[rank * len(MODELS):(rank + 1) * len(MODELS)], models_d))
    elif dom in ("fpfh.spfh", "fpfh.weight", "normals", "harris.response", "harris.nms"):
        r = p.fpfh_radius if dom.startswith("fpfh") else p.normal_radius
        per_k = {"fpfh.spfh": 32, "fpfh.weight": 16 + 132, "normals": 16, "harris.response": 32, "harris.nms": 16 + 4}[dom]
        per_q = {"fpfh.spfh": 16 + 132, "fpfh.weight": 16 + 132, "normals": 32, "harris.response": 16 + 4, "harris.nms": 16 + 1}[dom]
        alg = sum(km * per_k + nm * per_q + ks * per_k + ns * per_q for km, nm, ks, ns in neighbour_sums(r))
    elif dom == "icp.corr":
        alg = sum(cm.n * 32 for cm in models_d) * p.icp.max_iterations
    elif dom.startswith("grid.build"):
        # two grids (normals / FPFH radius) per cloud per registration: read 16 B, write 16 B sorted + 4 B key + 4 B slot per point
        alg = sum(2 * (cm.n + scene_d.n) * 40 for cm in models_d)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": peak_src, "launches_per_step": prof[dom][0], "ms_per_step": round(prof[dom][1], 4),
                "share_of_step": round(prof[dom][1] / step_ms, 4),
                "note": "configs[0..1] are launch-latency / L2 bound (working set < 1 MB, SURVEY 8d): the HBM fraction is reported, "
                        "not the criterion; the bandwidth-bound kernel is icp_1m.roofline"}
    tkey = "icp.corr@step(chair4 launch)" if dom == "icp.corr" else dom + "@chair4"
    if tkey in ncu_traffic:
        roofline["traffic"] = ncu_traffic[tkey]
        roofline["traffic_note"] = "dram bytes of the step's largest launch of this kernel (chair4), ncu --set full, profiles/r01_ncu_full_summary.md"
    if alg is not None and prof[dom][1] > 0:
        roofline["achieved"] = alg / (prof[dom][1] * 1e-3) / 1e9
        roofline["frac"] = roofline["achieved"] / hbm_peak
        roofline["algorithmic_bytes_per_step"] = alg

    log(f"profile pass done: top {top[0][0]} {top[0][1][1]:.3f} ms of {step_ms:.3f}")
    # ---- ICP iterations/s @ 1 M points (configs[2]), both directions
    icp_out = None
    if not args.no_icp and rank == 0:
        model, scan, gt = synth.icp_config(100_000, 1_000_000)
        cm, cs = api.Cloud(ctx, model), api.Cloud(ctx, scan)
        q = default_register_params()
        q.icp.max_iterations = 50
        q.icp.force_iterations = 1
        icp_out = {"workload": "configs[2]: synthetic 100k-point model vs 1M-point scan, 50 forced iterations, grid build included; "
                               "model_to_scan is the reference's direction (function.h:113-114) with PCL's unlimited correspondence "
                               "distance; scan_to_model (1M source points) caps it at 0.05 m"}
        for label, a, b, init, n_src, cap in (("model_to_scan", cm, cs, None, len(model), 0.0), ("scan_to_model", cs, cm, None, len(scan), 0.05)):
            q.icp.max_correspondence_distance = cap
            for _ in range(2):
                a.reset(); b.reset(); api.icp(a, b, q.icp, init)
            ms = 0.0
            reps = 5
            for _ in range(reps):
                a.reset(); b.reset()
                flush_buf.fill_(1); torch.cuda.synchronize()
                ctx.record(2); res = api.icp(a, b, q.icp, init); ctx.record(3)
                ms += ctx.elapsed_ms(2, 3)
            log(f"icp {label}: {ms / reps:.2f} ms per 50-iteration ICP")
            a.reset(); b.reset()
            ctx.profile_begin(); api.icp(a, b, q.icp, init); pr = ctx.profile_end()
            k_ms = pr["icp.corr"][1] / pr["icp.corr"][0]
            ach = n_src * 32 / (k_ms * 1e-3) / 1e9
            traffic = ncu_traffic.get("icp.corr@" + label)
            icp_out[label] = {"iters_per_s": 50 * reps / (ms * 1e-3), "ms_per_icp": ms / reps, "n_source": n_src,
                              "pose_err_vs_ground_truth": float(np.abs(res.matrix() - (gt if label == "model_to_scan" else np.linalg.inv(gt))).max()),
                              "fitness": float(res.fitness),
                              "roofline": {"kernel": "icp.corr", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                           "frac": ach / hbm_peak, "traffic": traffic, "avg_launch_ms": k_ms,
                                           "algorithmic_bytes_per_launch": n_src * 32, "peak_source": peak_src},
                              "kernel_ms": {k: round(v[1], 4) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:6]}}
        cm.free(); cs.free()

    # ---- the reference's own descriptor path (occupancy / TDF / 36-step yaw sweep / exhaustive consensus) on the pair main() loads
    native_out = None
    if rank == 0 and not args.no_native:
        try:
            from realtime_robot_b200.params import default_native_params
            nm_, ns_ = load_cloud("chair1"), to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "T0_m8111.pcd")))
            cmn, csn = api.Cloud(ctx, nm_), api.Cloud(ctx, ns_)
            npar = default_native_params()
            for _ in range(3):
                cmn.reset(); csn.reset(); rn = api.native_register(cmn, csn, npar)
            reps, ms = 20, 0.0
            for _ in range(reps):
                cmn.reset(); csn.reset()
                ctx.record(6); rn = api.native_register(cmn, csn, npar); ctx.record(7)
                ms += ctx.elapsed_ms(6, 7)
            npar0 = default_native_params(); npar0.use_plane_areas = 0
            ms0 = 0.0
            for _ in range(reps + 3):
                cmn.reset(); csn.reset()
                ctx.record(6); api.native_register(cmn, csn, npar0); ctx.record(7)
                ms0 += ctx.elapsed_ms(6, 7) if _ >= 3 else 0.0
            native_out = {"workload": "reference-native path, chair1.pcd (model) vs T0_m8111.pcd (scan) as main() loads them: getArea plane peel of "
                                      "both clouds, Harris, occupancy, TDF, 7 x 11 pair sweeps x 36 angles, screens, exhaustive consensus",
                          "registrations_per_s": reps / (ms * 1e-3),
                          "ms_per_registration": ms / reps, "ms_per_registration_without_plane_peel": ms0 / reps,
                          "keypoints": [int(rn.n_keypoints_src), int(rn.n_keypoints_tgt)],
                          "screened_pairs": int(rn.evaluated), "consensus": int(rn.inliers)}
            if world == 1 and not args.no_cpu_baseline:
                from oracle import orc
                orc.build(); orc.set_threads(1)
                t0 = time.time()
                mk = orc.harris3d(nm_, orc.normals(nm_, 0.05), 0.05, 0.01)[2]
                sk = orc.harris3d(ns_, orc.normals(ns_, 0.05), 0.05, 0.01)[2]
                orc.native_register(nm_, mk, ns_, sk, npar)
                native_out["cpu_port_1_thread_ms"] = 1e3 * (time.time() - t0)
            cmn.free(); csn.free()
        except Exception as e:          # the headline line must survive a failure of an auxiliary section
            native_out = {"error": repr(e)}
    # ---- configs[3]: synthetic 4 M-point scene, normals (r = .05) + FPFH (r = .08) of every point and at 100 000 keypoints.
    #      These are the HBM-scale gather kernels; their roofline uses SURVEY 8(d)'s algorithmic bytes (16 B per point and
    #      neighbour read, + 16 B per neighbour normal / 132 B per neighbour SPFH row, + the output row).
    scene_out = None
    if rank == 0 and not args.no_scene:
        try:
            n4m = 4_000_000
            side = max(4.0, (-12.0 + np.sqrt(144.0 + 4.0 * (n4m / 4700.0 * 0.8))) / 2.0)
            scene = synth.sample_rects(synth.room_rects((side, side, 3.0), n_boxes=max(4, int(side)), seed=synth.BASE_SEED), n4m, synth.BASE_SEED + 7)
            c4 = api.Cloud(ctx, scene)
            from realtime_robot_b200 import _lib
            L = _lib.lib()

            def run_normals():
                c4.reset(); L.rtr_normals(c4._h, 0.05, None)

            def run_fpfh():
                c4.reset(); L.rtr_normals(c4._h, 0.05, None); L.rtr_fpfh(c4._h, 0.08, None)

            def best_of(fn, reps):
                fn(); ms = []
                for _ in range(reps):
                    flush_buf.fill_(1); torch.cuda.synchronize()
                    ctx.record(2); fn(); ctx.record(3)
                    ms.append(ctx.elapsed_ms(2, 3))
                return min(ms)
            ms_n, ms_f = best_of(run_normals, 3), best_of(run_fpfh, 2)
            k5 = int(c4.radius_neighbors(0.05, counts_only=True)[0].sum())
            cnt8 = c4.radius_neighbors(0.08, counts_only=True)[0]
            k8 = int(cnt8.sum())
            ctx.profile_begin(); run_fpfh(); pr = ctx.profile_end()

            def roof(kernel, ms, alg):
                ach = alg / (ms * 1e-3) / 1e9
                return {"kernel": kernel, "bound": "hbm", "kernel_ms": round(ms, 3), "algorithmic_bytes": alg, "achieved": ach,
                        "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": peak_src}
            vox = np.floor(scene[:, :3] / 0.06).astype(np.int64)
            _, first = np.unique(vox[:, 0] * 1_000_003 + vox[:, 1] * 1009 + vox[:, 2], return_index=True)
            qidx = np.sort(np.random.default_rng(synth.BASE_SEED + 3).choice(first, min(100_000, len(first)), replace=False)).astype(np.int32)
            c4.fpfh_at(0.08, qidx)
            ctx.sync(); ctx.record(2); c4.fpfh_at(0.08, qidx); ctx.record(3)
            scene_out = {"workload": "configs[3]: synthetic 4M-point room (4700 pts/m2, 1 mm noise); grid build + normals r=.05; + FPFH r=.08 of every "
                                     "point; FPFH at 100 000 keypoints on the full surface (rtr_fpfh_at, host indices in, rows out)",
                         "points": n4m, "normals_ms_incl_grid": round(ms_n, 3), "normals_plus_fpfh_ms": round(ms_f, 3),
                         "neighbour_entries_r05": k5, "neighbour_entries_r08": k8,
                         "fpfh_at_100k_ms": round(ctx.elapsed_ms(2, 3), 3), "fpfh_at_neighbour_entries": int(cnt8[qidx].sum()),
                         "roofline_normals": roof("k_normals", pr["normals"][1], 16 * (n4m + k5) + 16 * n4m),
                         "roofline_spfh": roof("k_spfh", pr["fpfh.spfh"][1], 16 * (n4m + k8) + 16 * k8 + 132 * n4m),
                         "roofline_fpfh_weight": roof("k_fpfh_weight_tiled", pr["fpfh.weight"][1], 16 * (n4m + k8) + 132 * k8 + 132 * n4m)}
            c4.free()
            del scene
        except Exception as e:          # the headline line must survive a failure of an auxiliary section
            scene_out = {"error": repr(e)}

    # ---- PCD I/O either side of the path: 1 M points, the three DATA modes, file -> device cloud
    pcd_out = None
    if rank == 0 and not args.no_icp:
        try:
            big = synth.icp_config(1000, 1_000_000)[1]
            pcd_out = {"points": len(big)}
            with tempfile.TemporaryDirectory() as tmpdir:
                for mode, nm in ((api.PCD_ASCII, "ascii"), (api.PCD_BINARY, "binary"), (api.PCD_BINARY_COMPRESSED, "binary_compressed")):
                    f = os.path.join(tmpdir, nm + ".pcd")
                    t0 = time.perf_counter(); api.write_pcd(f, big, mode); tw = time.perf_counter() - t0
                    api.Cloud.from_pcd(ctx, f).free()
                    t0 = time.perf_counter(); cl = api.Cloud.from_pcd(ctx, f); ctx.sync(); tr = time.perf_counter() - t0
                    cl.free()
                    pcd_out[nm] = {"file_MB": round(os.path.getsize(f) / 1e6, 1), "write_ms": round(1e3 * tw, 1),
                                   "load_to_device_ms": round(1e3 * tr, 1), "load_Mpoints_per_s": round(len(big) / tr / 1e6, 1)}
        except Exception as e:
            pcd_out = {"error": repr(e)}
    log("cpu baseline")
    # ---- CPU baseline beside it (rank 0, N = 1): the oracle, one thread, the step's 8 registrations once
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import orc
        orc.build()
        orc.set_threads(1)
        t0 = time.time()
        n_done = 0
        for h in models_h:
            o = orc.register(h, scene_h, p)
            n_done += 1
            if time.time() - t0 > 40.0:
                break
        dt = time.time() - t0
        cpu = {"value": n_done / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"the step's first {n_done} of {len(MODELS)} registrations, once ({dt:.1f} s); the reference is single-threaded "
                         "(no /openmp in RealTimeRobot.vcxproj); PCL itself cannot be built here",
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        mine = recs[:len(MODELS)]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "repo .pcd clouds (xyz copies under data/clouds, written by tools/import_reference_clouds.py)",
                "config": workload_config(p),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "entry": "rtr_register_host (pinned host clouds)"},
                "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roofline, "kernel_share": kernel_share, "serialised_device_ms_per_step": round(step_ms, 3),
                "per_model_device_ms": per_model_ms, "per_model_latency_ms_alone": lat,
                "cpu_baseline": cpu, "icp_1m": icp_out, "scene_4m": scene_out, "native_path": native_out, "pcd_io_1m": pcd_out,
                "wall_ms_per_step_incl_l2_flush": 1e3 * wall_res / args.steps,
                "results": [{"model": m, "fitness": float(r.fitness), "inliers": int(r.inliers), "hypothesis": int(r.hypothesis),
                             "evaluated": int(r.evaluated), "converged": int(r.converged)} for m, r in zip(MODELS, mine)]}
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
