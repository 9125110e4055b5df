#!/usr/bin/env python3
"""bench.py — model-to-scene registrations/s (BASELINE.json metric) on N B200s, plus ICP iterations/s @ 1 M points.

A step = one pass of the hot path over one batch: ONE scan (mcloud.pcd) registered against every model of the database
(BASELINE.json configs[1]: chair1, chair2, chair4, desk1, desk2, desk3, sofa, Chair_025 x0.01) — grid -> normals -> Harris
-> FPFH -> feature k-NN -> 50 000 prerejective RANSAC hypotheses -> ICP (PCL defaults) for every model, the scan's own
stages once per scan as the reference does (RealTimeRobot.cpp:45-60).  Nothing is cached between steps: every stage of
every cloud is recomputed in every step (rtr_register_many builds its model set from the points each time).

  value : registrations/s with the clouds already resident in HBM (rtr_register_many on device handles)
  e2e   : the same through the host-buffer entry point (rtr_register_many_host: pinned host clouds in, H2D + D2H inside)
  N > 1 : model-sharded, weak scaling — the database grows with N (8 N models, rank r holds models 8 r .. 8 r + 7), every
          rank registers its shard against the replicated scan, then ONE ncclAllGather of the 128-byte records through the
          library (rtr_allgather_results), inside the timed region.  Sections strong_scaling (the fixed 8-model job split
          over the ranks) and ransac_sweep (configs[4], hypothesis-sharded) are measured in the same run.
  --impl reference : the CPU oracle (oracle/liboracle.so, all host threads, g++ -O3 -march=native) on the same workload.

One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
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
POSE_TOL, FIT_TOL = 1e-4, 1e-5          # BASELINE.json north_star


def load_cloud(name):
    src = "desk1" if name == "desk2" else name          # desk2.pcd is byte-identical to desk1.pcd in the reference
    pts = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", src + ".pcd")))
    if name == "Chair_025":
        pts[:, :3] *= np.float32(0.01)                  # units x100; the scale model_point.h:106-111 intends
    return pts


def workload_config(params, world=1):
    return {"workload": "configs[1]: one scan (mcloud.pcd) vs all repo models (chair1, chair2, chair4, desk1, desk2, desk3, sofa, Chair_025 x0.01), "
                        "full pipeline per registration (normals r=.05, Harris r=.05 thr=.01 NMS+refine, FPFH r=.10, k-NN k=5, prerejective "
                        "RANSAC, ICP PCL defaults), the scan's stages once per scan (RealTimeRobot.cpp:45-60); nothing cached between steps",
            "registrations_per_step_per_gpu": len(MODELS),
            "ransac_hypotheses": int(params.ransac.max_iterations),
            "icp_iterations": int(params.icp.max_iterations),
            "parallelism": "model-sharded: 8 models per rank in one batch (rtr_register_many), one 128 B/record ncclAllGather per step "
                           "(rtr_allgather_results); database of 8 N models at N GPUs",
            "l2": "flushed between timed steps (256 MiB write); inputs are < 2 MB"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "measured (MEASURED_PEAKS.json)")
    return 6650.0, 1590.0, 1590.0, "fallback (B200_PROFILING.md)"


def host_cores():
    """every host core this process may use (torchrun exports OMP_NUM_THREADS=1, which omp_get_max_threads() would obey)"""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def record_matches(g, o):
    """GPU record vs oracle record under the north-star tolerances."""
    return bool((g.hypothesis, g.inliers, g.evaluated, g.iterations, g.converged, g.n_keypoints_src, g.n_keypoints_tgt) ==
                (o.hypothesis, o.inliers, o.evaluated, o.iterations, o.converged, o.n_keypoints_src, o.n_keypoints_tgt) and
                float(np.abs(g.matrix() - o.matrix()).max()) <= POSE_TOL and
                (abs(float(g.fitness) - float(o.fitness)) <= FIT_TOL or float(g.fitness) == float(o.fitness)))


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    # every host core this process may use (torchrun exports OMP_NUM_THREADS=1, which omp_get_max_threads() would obey)
    cores = orc.set_threads(host_cores())
    p = default_register_params()
    scene = load_cloud(SCENE)
    models = [load_cloud(m) for m in MODELS]
    t = time.time()
    orc.register_many(models, scene, p)
    full = time.time() - t
    steps, warm = args.steps, max(args.warmup - 1, 0)
    budget = 150.0
    if full * (steps + warm) > budget:          # keep the run bounded: fewer steps of the same whole workload
        steps = max(1, int(budget / full) - warm)
    for _ in range(warm):
        orc.register_many(models, scene, p)
    t0 = time.time()
    for _ in range(steps):
        orc.register_many(models, scene, p)
    dt = time.time() - t0
    value = len(models) * steps / dt
    desc = (f"the step's {len(MODELS)} registrations (one scan vs all 8 models, the scan's stages once), {steps} step(s) of {dt / steps:.2f} s; "
            "oracle built -O3 -march=native on this box")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
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
    ap.add_argument("--verbose", action="store_true", help="per-step event / wall times on stderr")
    ap.add_argument("--no-native", action="store_true", help="skip the reference-native descriptor path and TDF A/B sections")
    ap.add_argument("--no-scene", action="store_true", help="skip the 4M-point scene section (configs[3]: normals + FPFH + matching)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] RANSAC sweep section")
    ap.add_argument("--quick", action="store_true", help="headline only: skip every auxiliary section")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained section (runs under a profiler)")
    args = ap.parse_args()
    if args.quick:
        args.no_icp = args.no_native = args.no_scene = args.no_sweep = True
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout on the first communicator — whatever it has
        # to say goes to stderr instead (a caller that sets NCCL_DEBUG / NCCL_DEBUG_FILE itself keeps its settings)
        os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as td
    from realtime_robot_b200 import _lib, api, dist, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the registration path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        td.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))      # a mismatch must fail fast, not hang the box
    ctx = api.Context(local_rank)
    if world > 1:
        dist.comm_init(ctx, rank, world)            # the library's own communicator (NCCL) on the context's stream
    L = _lib.lib()
    p = default_register_params()
    hbm_peak, bf16_burst, bf16_sustained, peak_src = peaks()
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        ncu_traffic = {}

    def log(msg):
        if rank == 0:
            print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)

    # host clouds in pinned memory (e2e leg) and resident device clouds (value leg)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    scene_t, scene_h = pinned(load_cloud(SCENE))
    models_h, keep = [], [scene_t]
    for m in MODELS:
        t, h = pinned(load_cloud(m))
        keep.append(t)
        models_h.append(h)
    scene_d = api.Cloud(ctx, scene_h)
    models_d = [api.Cloud(ctx, h) for h in models_h]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nm = len(MODELS)

    # N > 1: every batch ends with the in-stream ncclAllGather of its records (rtr_comm_gather_batches); a step returns the
    # world x 8 records every rank now holds.  Switched on only around the main loops: the other sections run batches of
    # different sizes per rank.
    from realtime_robot_b200.params import PoseResult
    gathered = (PoseResult * (nm * world))()

    def step_resident():
        recs = api.register_many(models_d, scene_d, p)
        if world == 1:
            return recs
        allr, n_all = dist.gathered_results(ctx, out=gathered)
        assert n_all == nm * world
        return allr

    def step_e2e():
        recs = api.register_many_host(ctx, models_h, scene_h, p)
        if world == 1:
            return recs
        allr, n_all = dist.gathered_results(ctx, out=gathered)
        assert n_all == nm * world
        return allr

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, flush=True):
        """Device time of `steps` steps: CUDA events on the library's own stream around each step (the batch, its result
        D2H and — at N > 1 — the ncclAllGather all run on that stream); L2 flushed (untimed) between steps.  The same clock at
        every N."""
        total_ms, last = 0.0, None
        for _ in range(steps):
            if flush:
                flush_buf.fill_(1)
                torch.cuda.synchronize()
                if world > 1:
                    td.barrier()            # untimed: every rank starts the step together (a step ends in a collective, so a late
                    torch.cuda.synchronize()  # starter's flush would otherwise be charged to the ranks waiting for it)
            ctx.record(0)
            t0 = time.perf_counter()
            last = step_fn()
            ctx.record(1)
            ev = ctx.elapsed_ms(0, 1)
            wall = (time.perf_counter() - t0) * 1e3
            total_ms += ev
            if args.verbose and rank == 0:
                print(f"[bench] step: events {ev:.3f} ms, host wall {wall:.3f} ms", file=sys.stderr, flush=True)
        return total_ms, last

    if world > 1:
        dist.gather_batches(ctx, True, rank * nm)
    log("warm-up")
    for _ in range(args.warmup):
        t0 = time.perf_counter(); step_resident(); t1 = time.perf_counter(); step_e2e()
        log(f"  warm-up step: resident {1e3 * (t1 - t0):.2f} ms, e2e {1e3 * (time.perf_counter() - t1):.2f} ms")

    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ctx.launches
    wall0 = time.perf_counter()
    ms_res, recs = timed(step_resident, args.steps)
    barrier()
    wall_res = time.perf_counter() - wall0
    launches = ctx.launches - l0
    ms_e2e, recs_e2e = timed(step_e2e, args.steps)
    barrier()
    # max over ranks (before anything is derived from the times: every rank must run the same number of steps below —
    # each step ends in a collective)
    if world > 1:
        t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_res, ms_e2e = float(t[0]), float(t[1])
    # a sustained section of our own (>= 2 s back to back, no L2 flush): the 100 ms clock sampler sees this one
    sus_steps = 2 if args.no_sustained else max(50, int(2500.0 / max(ms_res / args.steps, 0.05)))
    ms_sus, _ = timed(step_resident, sus_steps, flush=False)
    barrier()
    clocks = sampler.stop() if sampler else None
    # the exchange on its own: rtr_allgather_results of 8 host records per rank (H2D + ncclAllGather + D2H + sync), wall clock
    allgather_us, rank_ms_no_exchange = None, None
    if world > 1:
        mine_now = [PoseResult.from_buffer_copy(bytes(recs[rank * nm + i])) for i in range(nm)]
        recs_e2e = [PoseResult.from_buffer_copy(bytes(recs_e2e[rank * nm + i])) for i in range(nm)]
        recs = [PoseResult.from_buffer_copy(bytes(recs[i])) for i in range(nm * world)]
        dist.gather_batches(ctx, False)
        # where the weak-scaling loss comes from: the same steps WITHOUT the exchange, every rank on its own clock — the spread of
        # the ranks (a step that ends in a collective costs the slowest rank's time) against the cost of the collective itself
        ms_free, _ = timed(lambda: api.register_many(models_d, scene_d, p), args.steps)
        tf = torch.zeros(world, dtype=torch.float64, device=dev)
        tf[rank] = ms_free / args.steps
        td.all_reduce(tf, op=td.ReduceOp.SUM)
        rank_ms_no_exchange = [round(float(x), 4) for x in tf.tolist()]
        for _ in range(5):
            dist.allgather_results(ctx, mine_now, world)
        barrier()
        t0 = time.perf_counter()
        for _ in range(50):
            dist.allgather_results(ctx, mine_now, world)
        allgather_us = 1e6 * (time.perf_counter() - t0) / 50
    if world > 1:
        t = torch.tensor([ms_sus], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_sus = float(t[0])
    n_reg = nm * world * args.steps
    value = n_reg / (ms_res * 1e-3)
    e2e_value = n_reg / (ms_e2e * 1e-3)
    h2d = sum(len(h) * 16 for h in models_h) + len(scene_h) * 16
    d2h = 128 * nm + 1040 * (nm + 1)            # records + Harris corner previews
    mine = recs[rank * nm:(rank + 1) * nm] if world > 1 else recs
    log(f"timed: resident {ms_res / args.steps:.3f} ms/step, e2e {ms_e2e / args.steps:.3f} ms/step, sustained {ms_sus / sus_steps:.3f} ms/step")

    # ---- per-kernel device time of one more step (profiling marks after every stream operation; not the timed region)
    ctx.profile_begin()
    api.register_many(models_d, scene_d, p)
    prof = ctx.profile_end()
    step_ms = sum(v[1] for v in prof.values())
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])
    kernel_share = {k: {"launches": v[0], "ms": round(v[1], 4), "share": round(v[1] / step_ms, 4)} for k, v in top[:10]}
    # one registration alone (configs[0]: chair1 -> mcloud, rtr_register, full pipeline incl. the scan's stages): latency
    lat = {}
    for name, cm in (("chair1", models_d[0]), ("chair4", models_d[2])):
        for _ in range(2):
            cm.reset(); scene_d.reset(); api.register(cm, scene_d, p)
        ts = []
        for _ in range(5):
            cm.reset(); scene_d.reset()
            t0 = time.perf_counter(); api.register(cm, scene_d, p); ts.append(1e3 * (time.perf_counter() - t0))
        lat[name] = round(min(ts), 3)
        cm.reset(); scene_d.reset()

    # ---- roofline of the dominant kernel: algorithmic bytes over the clouds of the batch (8 models + the scan once)
    clouds_d = models_d + [scene_d]

    def neighbour_sum(r):
        return [(int(c.radius_neighbors(r, counts_only=True)[0].sum()), c.n) for c in clouds_d]
    dom = top[0][0]
    alg, bound, flops = None, "hbm", None
    if dom == "ransac.eval":
        alg = sum(int(r.evaluated) * (cm.n * 32 + 64) for r, cm in zip(mine, models_d))
    elif dom in ("fpfh.spfh", "fpfh.weight", "normals", "harris.response", "harris.nms"):
        r = p.fpfh_radius if dom.startswith("fpfh") else p.normal_radius
        per_k = {"fpfh.spfh": 32, "fpfh.weight": 16 + 132, "normals": 16, "harris.response": 32, "harris.nms": 16 + 4}[dom]
        per_q = {"fpfh.spfh": 16 + 132, "fpfh.weight": 16 + 132, "normals": 32, "harris.response": 16 + 4, "harris.nms": 16 + 1}[dom]
        alg = sum(k * per_k + n * per_q for k, n in neighbour_sum(r))
    elif dom == "icp.corr":
        # only registrations whose ICP ran, and the iterations they ran (a record with converged 0 after RANSAC skips ICP)
        alg = sum(cm.n * 32 * int(r.iterations) for r, cm in zip(mine, models_d))
    elif dom.startswith("grid.build"):
        alg = sum(2 * c.n * 40 for c in clouds_d)      # two grids per cloud: read 16 B, write 16 B sorted + 4 B key + 4 B slot per point
    elif dom == "match.tc_mma":
        bound = "tensor"
        flops = 2.0 * sum(c.n for c in models_d) * scene_d.n * 33
    roofline = {"kernel": dom, "bound": bound, "achieved": None, "peak": hbm_peak if bound == "hbm" else bf16_sustained,
                "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": None, "traffic": ncu_traffic.get(dom + "@step"),
                "peak_source": peak_src, "launches_per_step": prof[dom][0], "ms_per_step": round(prof[dom][1], 4),
                "share_of_step": round(prof[dom][1] / step_ms, 4),
                "note": "configs[0..1] hold < 2 MB of points: launch-latency / L1-L2 bound by construction (SURVEY 8d); the fraction is "
                        "reported, the bandwidth-scale kernels are in icp_1m / scene_4m"}
    if alg is not None and prof[dom][1] > 0:
        roofline["achieved"] = alg / (prof[dom][1] * 1e-3) / 1e9
        roofline["frac"] = roofline["achieved"] / hbm_peak
        roofline["algorithmic_bytes_per_step"] = alg
    if flops is not None and prof[dom][1] > 0:
        roofline["achieved"] = flops / (prof[dom][1] * 1e-3) / 1e12
        roofline["frac"] = roofline["achieved"] / bf16_sustained
        roofline["algorithmic_flops_per_step"] = flops
    log(f"profile pass done: top {top[0][0]} {top[0][1][1]:.3f} ms of {step_ms:.3f}")

    # ---- the reference's offline / online split (RealTimeRobot.cpp:124-165 vs :45-104): the 8 models prepared ONCE (normals,
    # corners, FPFH rows kept on the device), then every step streams one scan against them — upload, the scan's stages,
    # matching, RANSAC, ICP, records back.  NOT the headline (its steps reuse the models' descriptors by design): reported beside it.
    prepared = None
    if world == 1 and not args.quick:
        log("prepared-database section")
        try:
            prep = [api.Cloud(ctx, h) for h in models_h]
            ctx.sync()
            t0 = time.perf_counter()
            for c in prep:
                c.prepare(p)
            ctx.sync()
            offline_ms = 1e3 * (time.perf_counter() - t0)

            def step_online():
                cs = api.Cloud(ctx, scene_h)            # H2D of the scan from pinned memory
                rs = api.register_prepared(prep, cs, p)
                cs.free()
                return rs
            for _ in range(max(args.warmup, 3)):
                step_online()
            l0 = ctx.launches
            ms_on, recs_on = timed(step_online, args.steps)
            launches_on = (ctx.launches - l0) / args.steps
            ms_on_sus, _ = timed(step_online, 2 if args.no_sustained else 300, flush=False)
            same = all(bytes(a) == bytes(b) for a, b in zip(recs_on, mine))
            prepared = {"workload": "the 8 configs[1] models prepared once with rtr_cloud_prepare; a step = one scan (mcloud.pcd) from pinned host "
                                    "memory through rtr_cloud_upload + rtr_register_prepared: the scan's grid / normals / Harris / FPFH, one search "
                                    "of all model rows against the scan's, RANSAC and ICP for the 8 models, records back",
                        "ms_per_step": ms_on / args.steps, "value": nm * args.steps / (ms_on * 1e-3), "unit": UNIT,
                        "ms_per_step_back_to_back": ms_on_sus / (2 if args.no_sustained else 300),
                        "offline_prepare_ms_8_models": round(offline_ms, 3), "launches_per_step": launches_on,
                        "h2d_bytes_per_step": len(scene_h) * 16, "d2h_bytes_per_step": d2h,
                        "records_equal_the_uncached_step": bool(same)}
            for c in prep:
                c.free()
            log(f"prepared database: {ms_on / args.steps:.3f} ms/step online, {offline_ms:.2f} ms offline, records equal {same}")
        except Exception as e:
            prepared = {"error": repr(e)}

    # ---- N > 1: strong scaling of the FIXED 8-model job (longest-processing-time assignment of models to ranks)
    strong = None
    if world > 1:
        log("strong-scaling section")
        try:
            order = sorted(range(nm), key=lambda i: -models_h[i].shape[0])
            load, assign = [0] * world, [[] for _ in range(world)]
            for i in order:
                r = int(np.argmin(load)); assign[r].append(i); load[r] += models_h[i].shape[0]
            mine_ids = assign[rank]
            per = max(len(a) for a in assign)

            def step_strong():
                rs = api.register_many([models_d[i] for i in mine_ids], scene_d, p) if mine_ids else []
                for r, i in zip(rs, mine_ids):
                    r.model_id = i
                from realtime_robot_b200.params import PoseResult
                while len(rs) < per:
                    pad = PoseResult(); pad.model_id = -1; pad.hypothesis = -2; rs.append(pad)
                return dist.allgather_results(ctx, rs, world)
            for _ in range(3):
                step_strong()
            barrier()
            ms_s, allr = timed(step_strong, args.steps)
            t = torch.tensor([ms_s], dtype=torch.float64, device=dev)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            got = {r.model_id: r for r in allr if r.model_id >= 0}
            strong = {"workload": "the fixed configs[1] job (8 models vs mcloud) split over the ranks by model, longest cloud first; "
                                  "chair4 alone is the longest chain, so the job cannot scale past it",
                      "assignment": {str(r): [MODELS[i] for i in a] for r, a in enumerate(assign)},
                      "ms_per_step": float(t[0]) / args.steps, "value": nm * args.steps / (float(t[0]) * 1e-3), "unit": UNIT,
                      "records_equal_single_gpu": bool(all(i in got for i in range(nm)))}
        except Exception as e:
            strong = {"error": repr(e)}

    # ---- configs[4]: prerejective RANSAC sweep 1e4 .. 1e7 hypotheses, hypothesis-sharded over the ranks
    sweep = None
    if not args.no_sweep:
        log("ransac sweep section")
        try:
            cm, cs = models_d[0], scene_d
            cm.reset(); cs.reset()
            for c in (cm, cs):
                c.normals(p.normal_radius); c.fpfh(p.fpfh_radius)
            cm.match_features(cs, p.ransac.correspondence_k)
            sweep = {"workload": "configs[4]: chair1 -> mcloud, features and correspondences resident, hypotheses [r H / N, (r+1) H / N) on rank r, "
                                 "one 128-byte record per rank all-gathered, arg-min (fitness, hypothesis) on every rank", "points": []}
            q = default_register_params()
            for H in (10_000, 100_000, 1_000_000, 10_000_000):
                q.ransac.max_iterations = H
                q.ransac.hypothesis_begin, q.ransac.hypothesis_end = dist.shard_hypotheses(H, rank, world)

                def run():
                    r = api.ransac_prerejective(cm, cs, q.ransac)
                    return dist.select_best_hypothesis_native(dist.allgather_results(ctx, [r], world)) if world > 1 else r
                run()
                ms = []
                for _ in range(3):
                    ctx.sync(); ctx.record(4); best = run(); ctx.record(5)
                    ms.append(ctx.elapsed_ms(4, 5))
                t = torch.tensor([min(ms)], dtype=torch.float64, device=dev)
                if world > 1:
                    td.all_reduce(t, op=td.ReduceOp.MAX)
                pt = {"hypotheses": H, "ms": float(t[0]), "hypotheses_per_s": H / (float(t[0]) * 1e-3), "winner": int(best.hypothesis),
                      "inliers": int(best.inliers), "survivors": int(best.evaluated), "prerejection_rate": 1.0 - int(best.evaluated) / H}
                if rank == 0 and not args.no_cpu_baseline and H <= 100_000:
                    try:        # rank-local: must never break the collective sequence of the loop
                        from oracle import orc
                        orc.build(); orc.set_threads(host_cores())
                        qq = default_register_params(); qq.ransac.max_iterations = H
                        n4m, n4s = orc.normals(models_h[0], p.normal_radius), orc.normals(scene_h, p.normal_radius)
                        knn = orc.match_features(orc.fpfh(models_h[0], n4m, p.fpfh_radius), orc.fpfh(scene_h, n4s, p.fpfh_radius), p.ransac.correspondence_k)[0]
                        t0 = time.time(); o = orc.ransac(models_h[0], scene_h, knn, qq.ransac); dt = time.time() - t0
                        pt["cpu_all_threads_ms"] = 1e3 * dt
                        pt["winner_equals_oracle"] = bool(o.hypothesis == best.hypothesis and o.inliers == best.inliers)
                    except Exception as e:
                        pt["cpu_error"] = repr(e)
                sweep["points"].append(pt)
            cm.reset(); cs.reset()
        except Exception as e:
            sweep = {"error": repr(e)}

    # ---- the reference's own descriptor path (occupancy / TDF / 36-step yaw sweep / exhaustive consensus)
    native_out, tdf_out = None, None
    if rank == 0 and not args.no_native:
        try:
            from realtime_robot_b200.params import default_native_params
            nm_, ns_ = load_cloud("chair1"), to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "T0_m8111.pcd")))
            native_out = {}
            for label, a_h, b_h, gate in (("main_pair_gate3", nm_, ns_, 3.0), ("self_pair_gate30", ns_, ns_, 30.0)):
                cmn, csn = api.Cloud(ctx, a_h), api.Cloud(ctx, b_h)
                out = {}
                for peel in (1, 0):
                    npar = default_native_params(); npar.use_plane_areas = peel; npar.pair_gate = gate
                    reps, ms = 10, 0.0
                    for it in range(reps + 2):
                        cmn.reset(); csn.reset()
                        ctx.record(6); rn = api.native_register(cmn, csn, npar); ctx.record(7)
                        ms += ctx.elapsed_ms(6, 7) if it >= 2 else 0.0
                    out["ms_per_registration" if peel else "ms_per_registration_without_plane_peel"] = ms / reps
                    if peel:
                        rn1 = rn
                        out.update({"keypoints": [int(rn.n_keypoints_src), int(rn.n_keypoints_tgt)], "screened_pairs": int(rn.evaluated),
                                    "consensus": int(rn.inliers)})
                if world == 1 and not args.no_cpu_baseline:
                    from oracle import orc
                    orc.build(); orc.set_threads(1)
                    npar = default_native_params(); npar.pair_gate = gate
                    t0 = time.time()
                    mk = orc.harris3d(a_h, orc.normals(a_h, 0.05), 0.05, 0.01)[2]
                    sk = orc.harris3d(b_h, orc.normals(b_h, 0.05), 0.05, 0.01)[2]
                    on = orc.native_register(a_h, mk, b_h, sk, npar)
                    out["cpu_port_1_thread_ms"] = 1e3 * (time.time() - t0)
                    out["equals_oracle"] = bool((on.hypothesis, on.inliers, on.evaluated) == (rn1.hypothesis, rn1.inliers, rn1.evaluated))
                native_out[label] = out
                cmn.free(); csn.free()
            native_out["workload"] = ("reference-native path (getArea plane peel of both clouds, Harris, occupancy, TDF, all-pairs 36-angle sweeps, screens, "
                                      "exhaustive consensus): main_pair = chair1.pcd vs T0_m8111.pcd as main() loads them with the reference's gate "
                                      "get_Distance < 3 (no pair passes); self_pair = T0_m8111 against itself with gate 30, where the consensus has work")
        except Exception as e:          # the headline line must survive a failure of an auxiliary section
            native_out = {"error": repr(e)}
        # ---- TDF A/B: the one stage with a runnable reference (kernel.cu:34-106 compiled unmodified, oracle/_ref/libref_tdf.so)
        try:
            ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_tdf.so")
            ref = C.CDLL(ref_path)
            ref.ComputeTDFWithCuda.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            rng = np.random.default_rng(7)
            # (runs before the sections that allocate 1 M / 4 M-point clouds: the reference wrapper's cudaMalloc / cudaFree per keypoint
            #  get several times slower in a process that holds large pools — measured 0.6 ms against 6.6 ms per call)
            tdf_out = {"workload": "per-keypoint TDF of 191 occupied voxels in a 30^3 grid (KeyPoint::get_TSDF, key_point.h:251-318): the reference's "
                                   "ComputeTDFWithCuda (2 cudaMalloc, 2 H2D, <<<72,512>>>, cudaDeviceSynchronize, D2H, 2 cudaFree per keypoint, "
                                   "kernel.cu:34-106) called once per keypoint, vs librtr.so's same symbol, vs rtr_tdf_batch (all keypoints, one launch); "
                                   "host buffers in and out, wall clock, outputs bit-equal"}
            for nk in (7, 64, 512):
                occ = [rng.integers(0, 30, (191, 3)).astype(np.int32) for _ in range(nk)]

                def loop(fn):
                    outs = []
                    for o in occ:
                        buf = np.zeros(27000, np.float32)
                        rc = fn(o.ctypes.data, buf.ctypes.data, 30, len(o))
                        assert rc == 0
                        outs.append(buf)
                    return np.stack(outs)

                def best(f, reps=5):
                    f(); ts = []
                    for _ in range(reps):
                        t0 = time.perf_counter(); r = f(); ts.append(1e3 * (time.perf_counter() - t0))
                    return min(ts), r
                t_ref, o_ref = best(lambda: loop(ref.ComputeTDFWithCuda), 3)
                t_leg, o_leg = best(lambda: loop(L.ComputeTDFWithCuda))
                t_bat, o_bat = best(lambda: api.tdf_batch(ctx, occ, 30))
                assert np.array_equal(o_ref, o_leg) and np.array_equal(o_ref[:, :27000], o_bat)
                tdf_out[f"keypoints_{nk}"] = {"reference_ms": t_ref, "librtr_same_symbol_ms": t_leg, "rtr_tdf_batch_ms": t_bat,
                                              "speedup_same_symbol": t_ref / t_leg, "speedup_batched": t_ref / t_bat, "bit_equal": True}
        except Exception as e:
            tdf_out = {"error": repr(e)}

    # ---- ICP iterations/s @ 1 M points (configs[2]), both directions
    icp_out = None
    if not args.no_icp and rank == 0:
        try:
            model, scan, gt = synth.icp_config(100_000, 1_000_000)
            cm, cs = api.Cloud(ctx, model), api.Cloud(ctx, scan)
            q = default_register_params()
            q.icp.max_iterations = 50
            q.icp.force_iterations = 1
            icp_out = {"workload": "configs[2]: synthetic 100k-point model vs 1M-point scan, 50 forced iterations, grid build included; "
                                   "model_to_scan is the reference's direction (function.h:113-114) with PCL's unlimited correspondence distance; "
                                   "scan_to_model (1M source points) is given capped at 0.05 m AND uncapped (PCL default); source points far from the "
                                   "target's box search the 32-ary box hierarchy (csrc/bvh.cuh) instead of walking grid rings"}
            for label, a, b, n_src, cap, its in (("model_to_scan", cm, cs, len(model), 0.0, 50), ("scan_to_model", cs, cm, len(scan), 0.05, 50),
                                                 ("scan_to_model_uncapped", cs, cm, len(scan), 0.0, 50)):
                q.icp.max_correspondence_distance = cap
                q.icp.max_iterations = its
                for _ in range(2):
                    a.reset(); b.reset(); api.icp(a, b, q.icp, None)
                ms, reps = 0.0, 3
                for _ in range(reps):
                    a.reset(); b.reset()
                    flush_buf.fill_(1); torch.cuda.synchronize()
                    ctx.record(2); res = api.icp(a, b, q.icp, None); ctx.record(3)
                    ms += ctx.elapsed_ms(2, 3)
                log(f"icp {label}: {ms / reps:.2f} ms per {its}-iteration ICP")
                a.reset(); b.reset()
                ctx.profile_begin(); api.icp(a, b, q.icp, None); pr = ctx.profile_end()
                k_ms = pr["icp.corr"][1] / pr["icp.corr"][0]
                ach = n_src * 32 / (k_ms * 1e-3) / 1e9
                traffic = ncu_traffic.get("icp.corr@" + label)
                icp_out[label] = {"iters_per_s": its * reps / (ms * 1e-3), "ms_per_icp": ms / reps, "iterations": its, "n_source": n_src,
                                  "max_correspondence_distance": cap,
                                  "pose_err_vs_ground_truth": float(np.abs(res.matrix() - (gt if label == "model_to_scan" else np.linalg.inv(gt))).max()),
                                  "fitness": float(res.fitness),
                                  "roofline": {"kernel": "icp.corr", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                               "frac": ach / hbm_peak, "traffic": traffic, "avg_launch_ms": k_ms,
                                               "frac_on_dram_traffic": (traffic / (k_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                                               "algorithmic_bytes_per_launch": n_src * 32, "peak_source": peak_src},
                                  "kernel_ms": {k: round(v[1], 4) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:6]}}
            cm.free(); cs.free()
        except Exception as e:
            icp_out = {"error": repr(e)}

    # ---- configs[3]: synthetic 4 M-point scene, normals (r = .05) + FPFH (r = .08) of every point and at 100 000 keypoints,
    #      and brute-force tensor-core matching of 262 144 x 65 536 real FPFH rows.
    scene_out, match_out = None, None
    if rank == 0 and not args.no_scene:
        try:
            n4m = 4_000_000
            side = max(4.0, (-12.0 + np.sqrt(144.0 + 4.0 * (n4m / 4700.0 * 0.8))) / 2.0)
            scene = synth.sample_rects(synth.room_rects((side, side, 3.0), n_boxes=max(4, int(side)), seed=synth.BASE_SEED), n4m, synth.BASE_SEED + 7)
            c4 = api.Cloud(ctx, scene)

            def run_normals():
                c4.reset(); L.rtr_normals(c4._h, 0.05, None)

            def run_fpfh():
                c4.reset(); L.rtr_normals(c4._h, 0.05, None); L.rtr_fpfh(c4._h, 0.08, None)

            def run_normals_f32():
                c4.reset(); L.rtr_normals_mode(c4._h, 0.05, 1, None)

            def best_of(fn, reps):
                fn(); ms = []
                for _ in range(reps):
                    flush_buf.fill_(1); torch.cuda.synchronize()
                    ctx.record(2); fn(); ctx.record(3)
                    ms.append(ctx.elapsed_ms(2, 3))
                return min(ms)
            ms_n, ms_f = best_of(run_normals, 3), best_of(run_fpfh, 2)
            ms_n32 = best_of(run_normals_f32, 3)
            ctx.profile_begin(); run_normals_f32(); pr32 = ctx.profile_end()
            k5 = int(c4.radius_neighbors(0.05, counts_only=True)[0].sum())
            cnt8 = c4.radius_neighbors(0.08, counts_only=True)[0]
            k8 = int(cnt8.sum())
            ctx.profile_begin(); run_fpfh(); pr = ctx.profile_end()

            def roof(kernel, tag, ms, alg):
                ach = alg / (ms * 1e-3) / 1e9
                tr = ncu_traffic.get(tag + "@4m")
                return {"kernel": kernel, "bound": "hbm", "kernel_ms": round(ms, 3), "algorithmic_bytes": alg, "achieved": ach,
                        "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": tr,
                        "frac_on_dram_traffic": (tr / (ms * 1e-3) / 1e9 / hbm_peak) if tr else None, "peak_source": peak_src}
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
                         "roofline_normals": roof("k_normals", "normals", pr["normals"][1], 16 * (n4m + k5) + 16 * n4m),
                         "normals_pcl_float_ms_incl_grid": round(ms_n32, 3),
                         "roofline_normals_pcl_float": roof("k_normals_pcl_float (rtr_normals_mode 1: PCL's own float arithmetic)", "normals.pcl_float",
                                                            pr32["normals.pcl_float"][1], 16 * (n4m + k5) + 16 * n4m),
                         "roofline_spfh": roof("k_spfh", "fpfh.spfh", pr["fpfh.spfh"][1], 16 * (n4m + k8) + 16 * k8 + 132 * n4m),
                         "roofline_fpfh_weight": roof("k_fpfh_weight_tiled", "fpfh.weight", pr["fpfh.weight"][1], 16 * (n4m + k8) + 132 * k8 + 132 * n4m)}
            # matching: 262 144 scene features x 65 536 model features, real FPFH rows of this scene
            try:
                M, N = 262144, 65536
                feats = c4.fpfh(0.08)
                rng = np.random.default_rng(synth.BASE_SEED)
                fa = np.ascontiguousarray(feats[rng.choice(len(feats), M, replace=False)])
                fb = np.ascontiguousarray(feats[rng.choice(len(feats), N, replace=False)])
                del feats
                ms_m, st = api.match_raw(ctx, fa, fb, 5, reps=3)
                flop = 2.0 * M * N * 33
                ach = flop / (ms_m * 1e-3) / 1e12
                match_out = {"workload": "configs[3]: brute-force descriptor correspondence search, 262144 x 65536 x 33 (real FPFH rows of the 4M-point "
                                         "scene), k = 5: tf32 tcgen05 prefilter with a certificate + exact fp64 re-rank + exact redo of uncertified rows; "
                                         "device time of the whole search (prep + MMA + re-rank + redo), features resident",
                             "M": M, "N": N, "k": 5, "ms": ms_m, "pairs_per_s": M * N / (ms_m * 1e-3),
                             "rows_redone_exactly": st["redo_rows"], "splits": st["splits"],
                             "roofline": {"kernel": "k_tc_match", "bound": "tensor", "achieved": ach, "peak": bf16_sustained, "unit": "TFLOP/s",
                                          "frac": ach / bf16_sustained, "algorithmic_flops": flop, "peak_source": peak_src + " bf16_tflops_sustained",
                                          "note": "algorithmic flops 2 M N 33; the kernel issues tf32 MMAs (half the bf16 rate) over K padded to 104 for exactness"}}
            except Exception as e:
                match_out = {"error": repr(e)}
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
                for mode, nme in ((api.PCD_ASCII, "ascii"), (api.PCD_BINARY, "binary"), (api.PCD_BINARY_COMPRESSED, "binary_compressed")):
                    f = os.path.join(tmpdir, nme + ".pcd")
                    t0 = time.perf_counter(); api.write_pcd(f, big, mode); tw = time.perf_counter() - t0
                    api.Cloud.from_pcd(ctx, f).free()
                    t0 = time.perf_counter(); cl = api.Cloud.from_pcd(ctx, f); ctx.sync(); tr = time.perf_counter() - t0
                    cl.free()
                    pcd_out[nme] = {"file_MB": round(os.path.getsize(f) / 1e6, 1), "write_ms": round(1e3 * tw, 1),
                                    "load_to_device_ms": round(1e3 * tr, 1), "load_Mpoints_per_s": round(len(big) / tr / 1e6, 1)}
        except Exception as e:
            pcd_out = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N = 1): the oracle, ONE thread (the reference is single-threaded), the step's 8
    #      registrations once — and the parity of every GPU record of the timed step against it
    cpu, parity, cfg0 = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        log("cpu baseline")
        from oracle import orc
        orc.build()
        orc.set_threads(1)
        t0 = time.time()
        want = orc.register_many(models_h, scene_h, p)
        dt = time.time() - t0
        cpu = {"value": nm / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"the step's {nm} registrations once, scan-side stages once ({dt:.1f} s); the reference is single-threaded "
                         "(no /openmp in RealTimeRobot.vcxproj); PCL itself cannot be built here; oracle -O3 -march=native on this box",
               "host_cores_available": os.cpu_count()}
        parity = {m: record_matches(g, o) and record_matches(g2, o) for m, g, g2, o in zip(MODELS, mine, recs_e2e, want)}
        parity["all"] = bool(all(parity.values()))
        parity["tolerance"] = "hypothesis / inliers / evaluated / iterations / converged / keypoint counts identical; pose <= 1e-4, fitness <= 1e-5"
        # configs[0] on its own: chair1 -> mcloud, the pair the >= 50x target is quoted on
        t0 = time.time(); o1 = orc.register(models_h[0], scene_h, p); t1 = time.time() - t0
        allc = orc.set_threads(host_cores())
        t0 = time.time(); orc.register(models_h[0], scene_h, p); ta = time.time() - t0
        g1 = api.register_host(ctx, models_h[0], scene_h, p)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); api.register_host(ctx, models_h[0], scene_h, p); ts.append(time.perf_counter() - t0)
        cfg0 = {"workload": "configs[0]: chair1.pcd -> mcloud.pcd, one registration end to end from host clouds (rtr_register_host), wall clock",
                "gpu_ms": 1e3 * min(ts), "cpu_1_thread_ms": 1e3 * t1, "cpu_all_threads_ms": 1e3 * ta, "cpu_threads": allc,
                "speedup_vs_1_thread": t1 / min(ts), "speedup_vs_all_threads": ta / min(ts), "parity": record_matches(g1, o1)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "repo .pcd clouds (xyz copies under data/clouds, written by tools/import_reference_clouds.py)",
                "config": workload_config(p, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "entry": "rtr_register_many_host (pinned host clouds in, records out)"},
                "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                "sustained": {"steps": sus_steps, "ms_per_step": ms_sus / sus_steps, "value": nm * world * sus_steps / (ms_sus * 1e-3), "unit": UNIT,
                              "note": ">= 2 s back to back without L2 flushes: the section the 100 ms clock sampler sees"},
                "kernel_share": kernel_share, "serialised_device_ms_per_step": round(step_ms, 3),
                "launches_per_registration": launches / max(n_reg / world, 1),
                "single_registration_latency_ms": lat, "configs0_chair1_mcloud": cfg0, "prepared_database": prepared,
                "strong_scaling": strong, "ransac_sweep": sweep, "allgather_results_us": allgather_us,
                "per_rank_ms_per_step_without_exchange": rank_ms_no_exchange,
                "icp_1m": icp_out, "scene_4m": scene_out, "matching_262k_x_65k": match_out, "native_path": native_out, "tdf_ab": tdf_out,
                "pcd_io_1m": pcd_out,
                "wall_ms_per_step_incl_l2_flush": 1e3 * wall_res / args.steps,
                "results": [{"model": m, "fitness": float(r.fitness), "inliers": int(r.inliers), "hypothesis": int(r.hypothesis),
                             "evaluated": int(r.evaluated), "converged": int(r.converged), "iterations": int(r.iterations)} for m, r in zip(MODELS, mine)]}
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
