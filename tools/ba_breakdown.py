"""Per-kernel CUDA-event breakdown of one BundleAdjust-equivalent call + dense-solver task trace (torch-free).
usage: ba_breakdown.py [cfg] [lm_iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcptam_b200 import synth, capi

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prob = synth.make_ba_config(cfg, 0)
g = capi.BaHandle()
t = time.perf_counter(); g.load(prob); print(cfg, "load ms", 1e3 * (time.perf_counter() - t))
for _ in range(3):
    g.reset_state(); rc, st = g.compute(iters)
print(cfg, "rc", rc, "trials", st.total_trials, "launches", st.kernel_launches, "gpu_ms", st.gpu_ms, "it/s", 1e3 * rc / st.gpu_ms)
g.set_profiling(True)
g.reset_state(); rc, st = g.compute(iters)
tm = g.timing(); g.set_profiling(False)
print(cfg, "profiled (no speculation):", {k: round(v, 3) if isinstance(v, float) else v for k, v in tm.items()})
g.reset_state(); g.compute(3); g.reset_state()
g.solve_trace(arm=True)
d, s, r = g.lm_step(100.0)
tr = g.solve_trace()
T = (6 * g.n_pose_var + 31) // 32; nt = sum(2 + max(0, T - j - 2) for j in range(T))
tt = tr[:nt]; ch = tr[nt + 1: nt + 1 + T]; t0 = min(tt[:, 2].min(), ch[:, 2].min())
print(cfg, "solve: T", T, "worker tasks", nt, "factorisation span us", (max(tt[:, 4].max(), ch[:, 4].max()) - t0) / 1e3, "backsolve end us", (tr[nt, 0] - t0) / 1e3)
for j in range(T):
    if j < 4 or j >= T - 2:
        print(f"  chain step {j:2d}: inputs ready {(ch[j][2]-t0)/1e3:7.1f}  sub-diagonal solved {(ch[j][3]-t0)/1e3:7.1f}  factor published {(ch[j][4]-t0)/1e3:7.1f}")
