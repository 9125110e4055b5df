"""N > 1 host logic on CPU: world_size-2 gloo run of the shard planner and the 128-byte record all-gather."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from realtime_robot_b200 import dist
from realtime_robot_b200.params import PoseResult


def test_shard_planners():
    assert dist.shard_models(8, 0, 1) == list(range(8))
    got = sorted(sum((dist.shard_models(8, r, 3) for r in range(3)), []))
    assert got == list(range(8))
    for world in (1, 2, 4, 8, 3):
        spans = [dist.shard_hypotheses(50000, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 50000
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_record_roundtrip_and_selection():
    rs = []
    for i, (fit, h, conv) in enumerate([(0.5, 7, 1), (0.25, 9, 1), (0.25, 3, 1), (0.1, 1, 0)]):
        r = PoseResult(); r.fitness = fit; r.hypothesis = h; r.converged = conv; r.model_id = i
        r.pose[0] = float(i)
        rs.append(r)
    back = dist.bytes_to_records(dist.records_to_bytes(rs))
    assert [bytes(a) == bytes(b) for a, b in zip(rs, back)] == [True] * 4
    assert dist.select_best_hypothesis(rs).hypothesis == 3          # ties on fitness -> lowest hypothesis id
    assert dist.select_best_model(rs).model_id == 1
    assert dist.select_best_hypothesis([rs[3]]) is None


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch.distributed as td
    from realtime_robot_b200 import dist
    from realtime_robot_b200.params import PoseResult
    td.init_process_group("gloo")
    rank, world = td.get_rank(), td.get_world_size()
    mine = []
    for m in dist.shard_models(5, rank, world):
        r = PoseResult(); r.model_id = m; r.fitness = 1.0 / (1 + m); r.converged = 1; r.hypothesis = 100 + m
        mine.append(r)
    per = (5 + world - 1) // world
    allr = dist.all_gather_records(mine, per)
    ids = sorted(r.model_id for r in allr)
    best = dist.select_best_model(allr)
    b, e = dist.shard_hypotheses(1000, rank, world)
    print("RANK", rank, ids, best.model_id, b, e, flush=True)
    td.destroy_process_group()
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = sorted(l for o in outs for l in o.splitlines() if l.startswith("RANK"))
    assert lines[0] == "RANK 0 [0, 1, 2, 3, 4] 4 0 500"
    assert lines[1] == "RANK 1 [0, 1, 2, 3, 4] 4 500 1000"


SHARD_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as td
    from oracle import orc
    from realtime_robot_b200 import dist
    from realtime_robot_b200.params import default_register_params
    from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1
    td.init_process_group("gloo")
    rank, world = td.get_rank(), td.get_world_size()
    orc.set_threads(2)
    m = to_xyz1(read_pcd_xyz(os.path.join(%r, "data", "clouds", "chair1.pcd")))
    s = to_xyz1(read_pcd_xyz(os.path.join(%r, "data", "clouds", "mcloud.pcd")))
    p = default_register_params()
    fm, fs = orc.fpfh(m, orc.normals(m, 0.05), 0.10), orc.fpfh(s, orc.normals(s, 0.05), 0.10)
    knn, _ = orc.match_features(fm, fs, 5)
    H = 4000
    p.ransac.max_iterations = H
    full = orc.ransac(m, s, knn, p.ransac)
    p.ransac.hypothesis_begin, p.ransac.hypothesis_end = dist.shard_hypotheses(H, rank, world)
    mine = orc.ransac(m, s, knn, p.ransac)
    best = dist.select_best_hypothesis(dist.all_gather_records([mine], 1))
    same = (best.hypothesis, best.inliers) == (full.hypothesis, full.inliers) and bytes(best.pose) == bytes(full.pose) and best.fitness == full.fitness
    print("RANK", rank, int(best.hypothesis), int(same), flush=True)
    td.destroy_process_group()
""")


def test_world_size_2_gloo_hypothesis_sharded_ransac(tmp_path):
    """The multi-GPU RANSAC algorithm end to end on CPU: each rank evaluates its hypothesis range (the oracle stands in for
    the GPU), one all-gather of the 128-byte records, arg-min over (fitness, hypothesis) -> the unsharded winner on every rank."""
    script = tmp_path / "w2.py"
    script.write_text(SHARD_WORKER % (ROOT, ROOT, ROOT))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = sorted(l for o in outs for l in o.splitlines() if l.startswith("RANK"))
    assert len(lines) == 2 and lines[0].split()[2] == lines[1].split()[2] and int(lines[0].split()[2]) >= 0
    assert lines[0].split()[3] == "1" and lines[1].split()[3] == "1", lines


def test_native_hypothesis_selection_matches_python_rule():
    """rtr_select_best_hypothesis (host code of librtr.so, no GPU needed) == dist.select_best_hypothesis on random shard
    records incl. fitness ties, rejected shards and the nothing-accepted case."""
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(1, 9))
        rs = []
        for i in range(n):
            r = PoseResult()
            r.fitness = float(rng.choice([0.25, 0.5, 0.125, 1.0]))
            r.hypothesis = int(rng.integers(-1, 50))
            r.converged = int(rng.integers(0, 2)) if r.hypothesis >= 0 else 0
            r.evaluated = int(rng.integers(0, 100))
            r.inliers = int(rng.integers(0, 1000))
            r.pose[12] = float(i)
            rs.append(r)
        want = dist.select_best_hypothesis(rs)
        got = dist.select_best_hypothesis_native(rs)
        assert got.evaluated == sum(r.evaluated for r in rs)
        if want is None:
            assert got.hypothesis == -1 and got.converged == 0
        else:
            assert (got.hypothesis, got.fitness, got.pose[12], got.inliers) == (want.hypothesis, want.fitness, want.pose[12], want.inliers)


def test_allgather_without_communicator_is_a_copy_symbol_level():
    """single-GPU processes never touch NCCL: the library exports the comm entry points and loads without NCCL bound."""
    from realtime_robot_b200 import _lib
    L = _lib.lib()
    for s in ("rtr_comm_unique_id", "rtr_comm_init", "rtr_comm_destroy", "rtr_allgather_results", "rtr_select_best_hypothesis"):
        assert hasattr(L, s)
