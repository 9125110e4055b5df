#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle diagnostic (development aid; the graded checks are tests/ -m gpu).
Runs on the GPU box: reads only data/clouds and the repo's own libraries."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from realtime_robot_b200 import api, synth  # noqa: E402
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1  # noqa: E402


def cloud(name):
    return to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", name + ".pcd")))


def stage(name):
    def deco(fn):
        def run(*a, **k):
            t = time.time()
            try:
                fn(*a, **k)
                print(f"[ok]   {name}  ({time.time() - t:.2f}s)", flush=True)
            except Exception:
                print(f"[FAIL] {name}", flush=True)
                traceback.print_exc()
        return run
    return deco


ctx = api.Context(0)


@stage("tdf vs oracle and reference kernel.cu")
def check_tdf():
    rng = np.random.default_rng(1)
    ref = None
    p = os.path.join(ROOT, "oracle", "_ref", "libref_tdf.so")
    if os.path.exists(p):
        ref = C.CDLL(p)
    for n_occ, dim in [(0, 30), (1, 30), (37, 30), (191, 30), (2500, 30), (50, 12)]:
        occ = rng.integers(0, dim + 1, size=(n_occ, 3)).astype(np.int32)
        a = np.zeros(27000, np.float32)
        rc = api.compute_tdf_with_cuda(occ, a, dim, n_occ)
        o = orc.tdf(occ, dim)
        assert rc == 0
        assert np.array_equal(a[:dim ** 3], o), (n_occ, dim, np.abs(a[:dim**3] - o).max())
        if ref is not None:
            b = np.zeros(27000, np.float32)
            occ_c = np.ascontiguousarray(occ if n_occ else np.zeros((1, 3), np.int32))
            rc = ref.ComputeTDFWithCuda(occ_c.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), dim, n_occ)
            print("   ref rc", rc, "n_occ", n_occ, "dim", dim, "equal full buffer:", np.array_equal(a, b))
            assert np.array_equal(a, b)
    lists = [rng.integers(0, 30, size=(k, 3)).astype(np.int32) for k in (5, 0, 191, 64)]
    out = api.tdf_batch(ctx, lists, 30)
    for i, l in enumerate(lists):
        assert np.array_equal(out[i], orc.tdf(l, 30)), i
    print("   rc for num_occ=-1:", api.compute_tdf_with_cuda(np.zeros((0, 3), np.int32), np.zeros(27000, np.float32), 30, -1))


@stage("radius neighbours + nearest")
def check_index(name, r):
    pts = cloud(name)
    c = api.Cloud(ctx, pts)
    cnt, off, idx = c.radius_neighbors(r)
    ocnt, ooff, oidx = orc.radius_neighbors(pts, r, 1)
    print(f"   {name}: total {cnt.sum()} vs {ocnt.sum()}; counts equal {np.array_equal(cnt, ocnt)}; indices equal {np.array_equal(idx, oidx)}")
    assert np.array_equal(cnt, ocnt) and np.array_equal(idx, oidx)
    q = synth.apply(synth.rigid(3, -2, 10, (0.03, 0.02, -0.01)), pts)
    gi, gd = c.nearest(q)
    oi, od = orc.nearest(pts, q, 1)
    print(f"   nearest: idx mismatches {np.sum(gi != oi)}, d2 max abs diff {np.abs(gd - od).max()}")
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    far = q.copy(); far[:, :3] += 5.0
    gi, gd = c.nearest(far[:200]); oi, od = orc.nearest(pts, far[:200], 0)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    c.free()


def cmp(name, a, b, tol):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    bad = np.sum(~(d <= tol) & ~(np.isnan(a) & np.isnan(b)))
    print(f"   {name}: max abs diff {np.nanmax(d) if d.size else 0:.3e}, elements over {tol:g}: {bad} / {d.size}, bit-equal {np.mean(a.view(np.uint32) == b.view(np.uint32)) if a.size else 1:.6f}")
    return bad


@stage("per-cloud stages")
def check_features(name):
    pts = cloud(name)
    c = api.Cloud(ctx, pts)
    n4 = c.normals(0.05); o4 = orc.normals(pts, 0.05)
    cmp(f"{name} normals", n4, o4, 1e-5)
    resp, ki, kx = c.harris3d(0.05, 0.01)
    oresp, oki, okx = orc.harris3d(pts, o4, 0.05, 0.01)
    cmp(f"{name} harris response", resp, oresp, 1e-6)
    print(f"   keypoints gpu {ki.tolist()[:20]} oracle {oki.tolist()[:20]} n={len(ki)}/{len(oki)}")
    if len(ki) == len(oki) and len(ki):
        cmp(f"{name} refined corners", kx, okx, 1e-4)
    f = c.fpfh(0.10); of = orc.fpfh(pts, o4, 0.10)
    cmp(f"{name} fpfh", f, of, 1e-3)
    c.free()
    return f, of


@stage("matching + ransac + icp + register (chair1 -> mcloud)")
def check_pose():
    m, s = cloud("chair1"), cloud("mcloud")
    cm, cs = api.Cloud(ctx, m), api.Cloud(ctx, s)
    p = api.default_register_params()
    for c in (cm, cs):
        c.normals(p.normal_radius); c.fpfh(p.fpfh_radius)
    om4, os4 = orc.normals(m, 0.05), orc.normals(s, 0.05)
    ofm, ofs = orc.fpfh(m, om4, 0.1), orc.fpfh(s, os4, 0.1)
    gi, gd = cm.match_features(cs, 5)
    # feed the oracle the GPU's own features so this compares the matcher alone
    fm, fs = cm.fpfh(0.1), cs.fpfh(0.1)
    oi, od = orc.match_features(fm, fs, 5)
    print(f"   match: idx mismatches {np.sum(gi != oi)} / {gi.size}; dist max diff {np.abs(gd - od).max():.3e}")
    oi2, _ = orc.match_features(ofm, ofs, 5)
    print(f"   match vs oracle-features chain: idx mismatches {np.sum(gi != oi2)} / {gi.size}")
    r = api.ransac_prerejective(cm, cs, p.ransac)
    o = orc.ransac(m, s, gi, p.ransac)
    print("   ransac gpu   ", r.hypothesis, r.inliers, r.fitness, r.evaluated)
    print("   ransac oracle", o.hypothesis, o.inliers, o.fitness, o.evaluated)
    print("   pose diff", np.abs(r.matrix() - o.matrix()).max())
    ri = api.icp(cm, cs, p.icp, r.matrix())
    oi_ = orc.icp(m, s, p.icp, o.matrix())
    print("   icp gpu   ", ri.iterations, ri.converged, ri.fitness, ri.inliers)
    print("   icp oracle", oi_.iterations, oi_.converged, oi_.fitness, oi_.inliers)
    print("   pose diff", np.abs(ri.matrix() - oi_.matrix()).max())
    t = time.time(); rr = api.register(cm, cs, p); t1 = time.time() - t
    t = time.time(); orr = orc.register(m, s, p); t2 = time.time() - t
    print("   register gpu   ", rr.as_dict())
    print("   register oracle", orr.as_dict())
    print(f"   register pose diff {np.abs(rr.matrix() - orr.matrix()).max():.3e}; gpu (cached stages) {t1*1e3:.2f} ms, oracle {t2*1e3:.1f} ms")
    for rep in range(3):
        t = time.time(); rh = api.register_host(ctx, m, s, p); t3 = time.time() - t
        print(f"   register_host {t3*1e3:.2f} ms  launches so far {ctx.launches}  pose diff vs oracle {np.abs(rh.matrix() - orr.matrix()).max():.3e}")


@stage("icp synthetic 20k -> 200k, 30 forced iterations")
def check_icp_synth():
    m, s, gt = synth.icp_config(20000, 200000)
    cm, cs = api.Cloud(ctx, m), api.Cloud(ctx, s)
    p = api.default_register_params()
    p.icp.max_iterations = 30; p.icp.force_iterations = 1
    r = api.icp(cm, cs, p.icp)
    t = time.time(); r = api.icp(cm, cs, p.icp); tg = time.time() - t
    t = time.time(); o = orc.icp(m, s, p.icp); to = time.time() - t
    print(f"   gpu {tg*1e3:.1f} ms ({30/tg:.0f} it/s) oracle {to*1e3:.0f} ms; iterations {r.iterations}/{o.iterations}; fitness {r.fitness} / {o.fitness}")
    print(f"   pose diff gpu-oracle {np.abs(r.matrix() - o.matrix()).max():.3e}; vs ground truth {np.abs(r.matrix() - gt).max():.3e}")


if __name__ == "__main__":
    check_tdf()
    for nm in ("chair1", "sofa"):
        check_index(nm, 0.05)
    for nm in ("chair1", "mcloud", "sofa", "chair4"):
        check_features(nm)
    check_pose()
    check_icp_synth()
    print("launches", ctx.launches)
