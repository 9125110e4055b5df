"""rtr_register_many — one scan against many database models (RealTimeRobot.cpp:45-104 per model, README.md:10) through the
model-set path (segmented grids, shared launches): every record must equal rtr_register's for that model BIT FOR BIT, and
with it the CPU oracle's (tests/test_gpu_parity.py::test_register_matches_oracle pins rtr_register against the oracle)."""
import numpy as np
import pytest

from realtime_robot_b200.params import default_register_params

pytestmark = pytest.mark.gpu

MODELS = ["chair1", "chair2", "chair4", "desk1", "desk2", "desk3", "sofa", "Chair_025"]


@pytest.fixture(scope="module")
def api(gpu_ctx):
    from realtime_robot_b200 import api as a
    return a


def load(clouds, name):
    m = clouds("desk1" if name == "desk2" else name).copy()
    if name == "Chair_025":
        m[:, :3] *= np.float32(0.01)
    return m


def core(r):
    """every field of the record except model_id, bit-exact (pose and fitness as raw bytes)"""
    return (r.hypothesis, r.inliers, r.evaluated, r.iterations, r.converged, r.n_keypoints_src, r.n_keypoints_tgt,
            np.array(r.pose, dtype=np.float32).tobytes(), np.float32(r.fitness).tobytes())


def test_register_many_equals_register_and_oracle_on_configs1(api, gpu_ctx, orc, clouds):
    """BASELINE.json configs[1] at the benched settings: 8 models vs mcloud, 50 000 hypotheses, 10 ICP iterations."""
    p = default_register_params()
    scene = clouds("mcloud")
    hosts = [load(clouds, m) for m in MODELS]
    cs = api.Cloud(gpu_ctx, scene)
    cms = [api.Cloud(gpu_ctx, h) for h in hosts]
    l0 = gpu_ctx.launches
    many = api.register_many(cms, cs, p)
    launches_many = gpu_ctx.launches - l0
    assert [r.model_id for r in many] == list(range(len(MODELS)))
    l0 = gpu_ctx.launches
    singles = []
    for c in cms:
        c.reset(); cs.reset()
        singles.append(api.register(c, cs, p))
    launches_single = gpu_ctx.launches - l0
    for name, a, b in zip(MODELS, many, singles):
        assert core(a) == core(b), name
    # the batch shares its launches: the scan's stages once, every stage of the 8 models in one launch
    assert launches_many * 5 < launches_single, (launches_many, launches_single)
    # host-buffer form, asynchronous halves, repeatability
    many_h = api.register_many_host(gpu_ctx, hosts, scene, p)
    api.register_many_host_begin(gpu_ctx, hosts, scene, p)
    many_a = api.register_many_end(gpu_ctx)
    for a, b, c in zip(many, many_h, many_a):
        assert bytes(a) == bytes(b) == bytes(c)
    # and the oracle, scan-side stages once (orc_register_many)
    orc.set_threads(0)
    want = orc.register_many(hosts, scene, p)
    for name, g, o in zip(MODELS, many, want):
        assert (g.hypothesis, g.inliers, g.evaluated, g.iterations, g.converged) == (o.hypothesis, o.inliers, o.evaluated, o.iterations, o.converged), name
        assert (g.n_keypoints_src, g.n_keypoints_tgt) == (o.n_keypoints_src, o.n_keypoints_tgt), name
        assert np.abs(g.matrix() - o.matrix()).max() <= 1e-4 and abs(g.fitness - o.fitness) <= 1e-5, name
    for c in cms + [cs]:
        c.free()


def test_register_many_keypoints_are_the_harris_corners(api, gpu_ctx, clouds):
    p = default_register_params()
    p.ransac.max_iterations = 2000
    names = ["chair1", "T0_m8111", "desk1"]
    hosts = [load(clouds, m) for m in names]
    scene = clouds("mcloud")
    api.register_many_host(gpu_ctx, hosts, scene, p)
    for member, pts in enumerate(hosts + [scene]):
        c = api.Cloud(gpu_ctx, pts)
        c.normals(p.normal_radius)
        _, _, kx = c.harris3d(p.harris_radius, p.harris_threshold, p.harris_nms, p.harris_refine)
        got, n = api.register_many_keypoints(gpu_ctx, member)
        assert n == len(kx) and np.array_equal(got, kx[:64]), member
        c.free()


def test_register_many_point_to_plane_and_hypothesis_shards(api, gpu_ctx, clouds):
    """estimator 1 (6x6 LLS on the scan's normals) and a hypothesis shard [begin, end) through the batch."""
    p = default_register_params()
    p.icp.estimator = 1
    p.ransac.max_iterations = 20000
    p.ransac.hypothesis_begin, p.ransac.hypothesis_end = 5000, 17000
    hosts = [load(clouds, m) for m in ("chair2", "chair1", "Chair_025")]
    scene = clouds("mcloud")
    many = api.register_many_host(gpu_ctx, hosts, scene, p)
    for h, r in zip(hosts, many):
        s = api.register_host(gpu_ctx, h, scene, p)
        s.model_id = r.model_id
        assert bytes(s) == bytes(r)
    assert any(r.converged for r in many)


def test_register_many_edge_cases(api, gpu_ctx, clouds):
    """a single model; duplicates; degenerate members (empty / 2 points: per-model fallback inside the call); > 31 models."""
    p = default_register_params()
    p.ransac.max_iterations = 3000
    scene = clouds("mcloud")
    c1 = load(clouds, "chair1")
    ref = api.register_host(gpu_ctx, c1, scene, p)
    one = api.register_many_host(gpu_ctx, [c1], scene, p)
    assert len(one) == 1 and bytes(one[0]) == bytes(ref)
    tiny = np.array([[0, 0, 0, 1], [0.01, 0, 0, 1]], np.float32)
    mixed = api.register_many_host(gpu_ctx, [c1, tiny, c1], scene, p)
    assert [r.model_id for r in mixed] == [0, 1, 2]
    assert mixed[1].converged == 0 and mixed[1].hypothesis == -1
    for r in (mixed[0], mixed[2]):
        r.model_id = 0
        assert bytes(r) == bytes(ref)
    many = api.register_many_host(gpu_ctx, [c1] * 33, scene, p)          # two batches: 31 + 2
    assert [r.model_id for r in many] == list(range(33))
    for r in many:
        r.model_id = 0
        assert bytes(r) == bytes(ref)
    # a scan that matches nothing: every record is the "nothing accepted" record
    far = scene.copy(); far[:, :3] *= np.float32(0.01)
    none = api.register_many_host(gpu_ctx, [c1, load(clouds, "desk1")], far, p)
    for h, r in zip([c1, load(clouds, "desk1")], none):
        s = api.register_host(gpu_ctx, h, far, p)
        s.model_id = r.model_id
        assert bytes(s) == bytes(r)
    # invalid parameters come back as RTR_ERR_INVALID
    from realtime_robot_b200 import _lib
    q = default_register_params(); q.fpfh_radius = 0.0
    with pytest.raises(_lib.RtrError) as e:
        api.register_many_host(gpu_ctx, [c1], scene, q)
    assert e.value.code == 1


def test_cpp_host_database_mode_matches_ctypes_path(api, gpu_ctx, clouds):
    """The C++ host mirror's registerModelsToScene (rtr_register_many_host behind pcl::PointCloud in / Eigen::Matrix4f out):
    same records as the ctypes path, and the corners come back as ModelPoint / ScanPoint::key_coordinates would hold them."""
    import os
    import re
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "realtime_robot_b200", "realtime_robot")
    data = os.path.join(ROOT, "data", "clouds")
    names = ["chair1", "chair2", "desk1"]
    r = subprocess.run([exe, "--database", os.path.join(data, "mcloud.pcd")] + [os.path.join(data, n + ".pcd") for n in names] + ["--hypotheses", "20000"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    p = default_register_params()
    p.ransac.max_iterations = 20000
    want = api.register_many_host(gpu_ctx, [clouds(n) for n in names], clouds("mcloud"), p)
    rows = re.findall(r"model (\d+) converged (\d) hypothesis (-?\d+) inliers (\d+) fitness (\S+) keypoints (\d+) scan_keypoints (\d+)", r.stdout)
    assert len(rows) == len(names), r.stdout
    for row, w in zip(rows, want):
        assert (int(row[1]) != 0, int(row[2]), int(row[3]), int(row[5]), int(row[6])) == (w.converged != 0, w.hypothesis, w.inliers, w.n_keypoints_src, w.n_keypoints_tgt)


def test_cpp_reference_signatures_run_mains_loops(api, gpu_ctx, clouds):
    """main()'s own loops (RealTimeRobot.cpp:49-104) through the reference's signatures — KeyPoint::getOccupiedGrid / get_TSDF,
    get_Distance(matrix, model_key, scan_key), match_by_*, Ransac(pairpoint, 50, cloud, mcloud) — give the pair count and the
    pose of the batched device path (rtr_native_register)."""
    import os
    import re
    import subprocess
    from conftest import ROOT
    from realtime_robot_b200.params import default_native_params
    exe = os.path.join(ROOT, "realtime_robot_b200", "realtime_robot")
    f = os.path.join(ROOT, "data", "clouds", "T0_m8111.pcd")
    r = subprocess.run([exe, f, f, "--reference-main", "--gate", "30"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    pairs = int(re.search(r"pairs (\d+)", r.stdout).group(1))
    nums = [float(v) for v in re.search(r"matrix\n(.*?)Running", r.stdout, re.S).group(1).split()]
    got = np.array(nums, dtype=np.float32).reshape(4, 4)
    c = clouds("T0_m8111")
    cm, cs = api.Cloud(gpu_ctx, c), api.Cloud(gpu_ctx, c)
    q = default_native_params(); q.pair_gate = 30.0
    w = api.native_register(cm, cs, q)
    assert pairs == w.evaluated and pairs > 0
    assert np.abs(got - w.matrix()).max() < 1e-4
    cm.free(); cs.free()


def test_register_prepared_is_the_online_half_of_register_many(api, gpu_ctx, clouds):
    """rtr_cloud_prepare + rtr_register_prepared — the reference's offline / online split (RealTimeRobot.cpp:124-165 vs :45-104):
    models prepared once, scans streamed against them.  Records, corner counts and corner previews equal rtr_register_many's bit for
    bit, for a fresh scan, for a prepared scan, for a second scan against the same prepared models, and through _begin / _end."""
    from realtime_robot_b200._lib import RtrError
    p = default_register_params()
    hosts = [load(clouds, m) for m in MODELS]
    scenes = [clouds("mcloud"), clouds("T0_m8111")]
    cms = [api.Cloud(gpu_ctx, h) for h in hosts]
    fresh = [api.Cloud(gpu_ctx, h) for h in hosts]           # never prepared: the reference batch
    # not prepared yet: refused, nothing computed silently
    with pytest.raises(RtrError):
        api.register_prepared(cms, api.Cloud(gpu_ctx, scenes[0]), p)
    for c in cms:
        c.prepare(p)
    for scene in scenes:
        cs = api.Cloud(gpu_ctx, scene)
        l0 = gpu_ctx.launches
        want = api.register_many(fresh, cs, p)
        launches_many = gpu_ctx.launches - l0
        want_kp = [api.register_many_keypoints(gpu_ctx, m) for m in range(len(MODELS) + 1)]
        cs.free()
        cs = api.Cloud(gpu_ctx, scene)                        # a fresh scan: its stages run inside the call
        l0 = gpu_ctx.launches
        got = api.register_prepared(cms, cs, p)
        launches = gpu_ctx.launches - l0
        got_kp = [api.register_many_keypoints(gpu_ctx, m) for m in range(len(MODELS) + 1)]
        for name, a, b in zip(MODELS, got, want):
            assert bytes(a) == bytes(b), name
        for (ka, na), (kb, nb) in zip(got_kp, want_kp):
            assert na == nb and np.array_equal(ka, kb)
        again = api.register_prepared(cms, cs, p)             # the scan is prepared now: nothing of it is recomputed
        api.register_prepared_begin(cms, cs, p)
        halves = api.register_many_end(gpu_ctx)
        for a, b, c in zip(got, again, halves):
            assert bytes(a) == bytes(b) == bytes(c)
        assert launches <= launches_many + 4, (launches, launches_many)     # the scan's stages + two gathers: no per-model launches
        cs.free()
    # other stage parameters than the models were prepared with: refused; reset() drops the preparation
    q = default_register_params()
    q.fpfh_radius = 0.08
    cs = api.Cloud(gpu_ctx, scenes[0])
    with pytest.raises(RtrError):
        api.register_prepared(cms, cs, q)
    cms[2].reset()
    with pytest.raises(RtrError):
        api.register_prepared(cms, cs, p)
    cms[2].prepare(p)
    got = api.register_prepared(cms[:3], cs, p)               # a subset of the database
    want = api.register_many(fresh[:3], api.Cloud(gpu_ctx, scenes[0]), p)
    for a, b in zip(got, want):
        assert bytes(a) == bytes(b)
    # degenerate members are refused by the prepared path (rtr_register_many falls back to rtr_register for them instead)
    tiny = api.Cloud(gpu_ctx, hosts[0][:2])
    tiny.prepare(p)
    with pytest.raises(RtrError):
        api.register_prepared([tiny], cs, p)
    empty = api.Cloud(gpu_ctx, np.zeros((0, 4), np.float32))
    with pytest.raises(RtrError):
        api.register_prepared(cms[:2], empty, p)
    with pytest.raises(RtrError):
        api.register_prepared([cms[0], cs], cs, p)           # the scan cannot be one of its own models
    for c in cms + fresh + [cs, tiny, empty]:
        c.free()


def test_cpp_host_model_database_offline_online(api, gpu_ctx, clouds):
    """The C++ host mirror's ModelDatabase (add = rtr_cloud_prepare, match = rtr_register_prepared) behind the Linux driver's
    --database-online mode: two scans against three prepared models give the records of the ctypes batch path."""
    import os
    import re
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "realtime_robot_b200", "realtime_robot")
    data = os.path.join(ROOT, "data", "clouds")
    names, scans = ["chair1", "chair2", "desk1"], ["mcloud", "T0_m8111"]
    r = subprocess.run([exe, "--database-online"] + [os.path.join(data, n + ".pcd") for n in names] + ["--scans"] +
                       [os.path.join(data, s + ".pcd") for s in scans] + ["--hypotheses", "20000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert re.search(r"time to preprocess 3 database models", r.stdout)
    p = default_register_params()
    p.ransac.max_iterations = 20000
    rows = re.findall(r"scan \S+ model (\d+) converged (\d) hypothesis (-?\d+) inliers (\d+) fitness (\S+)", r.stdout)
    assert len(rows) == len(names) * len(scans), r.stdout
    want = []
    for s in scans:
        want += api.register_many_host(gpu_ctx, [clouds(n) for n in names], clouds(s), p)
    for row, w in zip(rows, want):
        assert (int(row[1]) != 0, int(row[2]), int(row[3])) == (w.converged != 0, w.hypothesis, w.inliers)
        assert abs(float(row[4]) - w.fitness) <= 1e-5 * max(1e-3, abs(w.fitness))        # six printed digits


def test_ransac_block_lists_and_grid_walk_give_the_same_records(api, gpu_ctx):
    """The prerejective inlier test reads the scan either as flattened 3x3x3 block lists (default for batches and large sweeps) or by
    walking the grid's 9 row ranges (RTR_RANSAC_BLOCKLISTS=0): the candidate set of a query is the same, so are the records — a batch
    of three models and a 250 000-hypothesis single registration, bit for bit."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = ("import sys, json; sys.path.insert(0, %r)\n"
            "import os\n"
            "from realtime_robot_b200 import api\n"
            "from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1\n"
            "load = lambda n: to_xyz1(read_pcd_xyz(os.path.join(%r, 'data', 'clouds', n + '.pcd')))\n"
            "ctx = api.Context(0)\n"
            "p = api.default_register_params()\n"
            "p.ransac.max_iterations = 20000\n"
            "rs = api.register_many_host(ctx, [load('chair1'), load('chair4'), load('desk1')], load('mcloud'), p)\n"
            "p.ransac.max_iterations = 250000\n"
            "rs.append(api.register_host(ctx, load('chair2'), load('mcloud'), p))\n"
            "print(json.dumps([bytes(r).hex() for r in rs]))\n") % (ROOT, ROOT)
    outs = []
    for mode in ("1", "0"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RTR_RANSAC_BLOCKLISTS=mode), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1] and len(outs[0]) == 4
