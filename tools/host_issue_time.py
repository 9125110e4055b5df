import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from realtime_robot_b200 import api
from realtime_robot_b200.pcd import read_pcd_xyz, to_xyz1
ROOT='/root/repo'
ctx = api.Context(0)
s_h = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", "mcloud.pcd")))
p = api.default_register_params()
for name in ["chair1", "chair4", "sofa"]:
    m_h = to_xyz1(read_pcd_xyz(os.path.join(ROOT, "data", "clouds", name + ".pcd")))
    m, s = api.Cloud(ctx, m_h), api.Cloud(ctx, s_h)
    for _ in range(3):
        m.reset(); s.reset(); api.register(m, s, p)
    tb = te = tr = 0
    for _ in range(20):
        t0 = time.perf_counter(); m.reset(); s.reset(); t1 = time.perf_counter()
        api.register_begin(m, s, p); t2 = time.perf_counter()
        api.register_end(ctx); t3 = time.perf_counter()
        tr += t1 - t0; tb += t2 - t1; te += t3 - t2
    l0 = ctx.launches; m.reset(); s.reset(); api.register(m, s, p); l1 = ctx.launches
    print(f"{name}: reset {tr/20*1e6:.0f} us, begin {tb/20*1e6:.0f} us, end(wait) {te/20*1e6:.0f} us, counted launches {l1-l0}")
