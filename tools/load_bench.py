"""mcp_ba_load cost (host marshalling + uploads + device list construction) vs the number of marshalling threads.
usage: load_bench.py [cfg]   (MCP_BA_HOST_THREADS is read per load)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcptam_b200 import synth, capi

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
prob = synth.make_ba_config(cfg, 0)
for thr in (1, 2, 4, 8, 12, 16):
    os.environ["MCP_BA_HOST_THREADS"] = str(thr)
    g = capi.BaHandle()
    for _ in range(5):
        g.load(prob)
    t = time.perf_counter()
    n = 30
    for _ in range(n):
        g.load(prob)
    print(cfg, "threads", thr, "load ms", 1e3 * (time.perf_counter() - t) / n, flush=True)
    g.close()
