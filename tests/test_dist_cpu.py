"""World-size-2 gloo test (CPU) of the multi-GPU decomposition: map points are partitioned with the product's
own partition function, every rank assembles the Schur-reduced camera system of its shard (from the oracle's
Jacobians), the shards are all-reduced, and the replicated dense solve reproduces the single-process LM step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _shard_system(prob, o, p_lo, p_hi, lam, sigma_lim_sq):
    """Schur-reduced system of points [p_lo, p_hi) from the oracle's residuals/Jacobians (numpy)."""
    npv = int((prob.pose_fixed == 0).sum())
    var = -np.ones(prob.n_pose, int); var[prob.pose_fixed == 0] = np.arange(npv)
    S = np.zeros((6 * npv, 6 * npv)); r = np.zeros(6 * npv)
    e, chi2 = o.eval()
    sig = np.sqrt(sigma_lim_sq)
    for p in range(p_lo, p_hi):
        ms = np.flatnonzero(prob.meas_pt == p)
        V = np.zeros((3, 3)); gp = np.zeros(3); W = {}
        for m in ms:
            jo, js, jp = o.jacobians(m)
            info = 1.0 / np.sqrt(prob.meas_noise[m])
            w = info * (1.0 if chi2[m] <= sigma_lim_sq else sig / np.sqrt(chi2[m]))
            Js = {}
            for pid, J in ((prob.meas_chain[m][0], jo[0]), (prob.pt_chain[p][0], js[0])):
                if var[pid] >= 0 and np.abs(J).sum() > 0:
                    Js[var[pid]] = Js.get(var[pid], 0) + J
            for a, Ja in Js.items():
                r[6 * a:6 * a + 6] -= w * Ja.T @ e[m]
                for b, Jb in Js.items():
                    S[6 * a:6 * a + 6, 6 * b:6 * b + 6] += w * Ja.T @ Jb
                W[a] = W.get(a, 0) + w * Ja.T @ jp
            V += w * jp.T @ jp; gp -= w * jp.T @ e[m]
        Vi = np.linalg.inv(V + lam * np.eye(3))
        for a, Wa in W.items():
            r[6 * a:6 * a + 6] -= Wa @ Vi @ gp
            for b, Wb in W.items():
                S[6 * a:6 * a + 6, 6 * b:6 * b + 6] -= Wa @ Vi @ Wb.T
    return S, r


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcptam_b200 import capi, synth
    from oracle.oracle import OracleBA
    prob = synth.make_ba_config("tiny", seed=1)
    o = OracleBA(prob)
    lam = 5.0
    rc, d_ref, sig_raw, _ = o.lm_step(lam, -1.0, 0)
    sig_lim = max(sig_raw, 0.25)
    part = capi.ba_partition(prob.n_pt, prob.meas_pt, world)
    S, r = _shard_system(prob, o, int(part[rank]), int(part[rank + 1]), lam, sig_lim)
    buf = torch.from_numpy(np.concatenate([S.ravel(), r]))
    dist.all_reduce(buf)                                            # the NCCL allreduce of the product
    n = len(r)
    S = buf[: n * n].numpy().reshape(n, n) + lam * np.eye(n); r = buf[n * n:].numpy()
    dc = np.linalg.solve(S, r)                                      # replicated dense solve
    err = np.linalg.norm(dc - d_ref[:n]) / np.linalg.norm(d_ref[:n])
    q.put((rank, float(err), [int(x) for x in part]))
    dist.destroy_process_group()


def test_point_sharded_schur_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    parts = {tuple(r[2]) for r in res}
    assert len(parts) == 1                                          # every rank derives the same partition
    for _, err, _ in res:
        assert err < 1e-8
