"""End-to-end BundleAdjust-equivalent call (host arrays -> mcp_ba_load -> compute -> read back), torch-free.
usage: e2e_bench.py [cfg] [lm_iters] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcptam_b200 import synth, capi

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
prob = synth.make_ba_config(cfg, 0)
g = capi.BaHandle()
parts = np.zeros(3); n_it = 0
for s in range(5 + reps):
    t0 = time.perf_counter(); g.load(prob)
    t1 = time.perf_counter(); rc, st = g.compute(iters)
    t2 = time.perf_counter(); P, X = g.poses(), g.points(); _ = g.outliers()
    t3 = time.perf_counter()
    if s >= 5:
        parts += (t1 - t0, t2 - t1, t3 - t2); n_it += rc
print(cfg, "e2e LM it/s", n_it / parts.sum(), "ms per call: load %.3f compute %.3f read-back %.3f" % tuple(1e3 * parts / reps), "gpu_ms(loop)", st.gpu_ms)
