"""Host marshalling of mcp_ba_load (mcptam_b200/csrc/ba_prep.hpp) checked on the CPU against a definition-level
Python restatement: measurement order, per-point pose-slot lists (PoseChainHelper::MoveTogether semantics,
src/ChainBundle.cc:151-170), visiting order, pose-block work lists, multi-rank partition and the argument errors.
The same header is compiled into the CUDA library; this runs it through the g++-built shim (no GPU needed)."""
import copy

import numpy as np
import pytest

from mcptam_b200 import capi, synth


def _restate(prob, rank, world):
    n_pt, n_meas = prob.n_pt, prob.n_meas
    pose_var = np.full(prob.n_pose, -1, np.int64)
    mov = np.flatnonzero(np.asarray(prob.pose_fixed) == 0)
    pose_var[mov] = np.arange(len(mov))
    npv = len(mov)
    meas_pt = np.asarray(prob.meas_pt)
    order = np.argsort(meas_pt, kind="stable")
    off = np.concatenate([[0], np.cumsum(np.bincount(meas_pt, minlength=n_pt))]).astype(np.int64)
    slot_var, slot_pt, slot_off = [], [], [0]
    meas_b = np.zeros((n_meas, 4), np.int64)
    pt_info = np.zeros((n_pt, 4), np.int64)
    for p in range(n_pt):
        src0 = int(prob.pt_chain[p][0])
        sv = int(pose_var[src0])
        movable = not prob.pt_fixed[p]
        seen, any_src = set(), False
        rows = []
        for q in range(off[p], off[p + 1]):
            m = order[q]
            obs0 = int(prob.meas_chain[m][0])
            has_jac = obs0 != src0
            ov = int(pose_var[obs0]) if has_jac else -1
            has_src = has_jac and sv >= 0
            any_src |= has_src
            if ov >= 0:
                seen.add(ov)
            rows.append((q, ov, has_src))
        slots = []
        if movable:
            if any_src:
                seen.add(sv)
            slots = sorted(seen)
        for q, ov, has_src in rows:
            meas_b[q] = (ov, slots.index(ov) if (movable and ov >= 0) else -1, int(has_src), p)
        pt_info[p] = (src0, int(prob.pt_chain[p][1]), sv, slots.index(sv) if (movable and any_src) else -1)
        slot_var += slots
        slot_pt += [p] * len(slots)
        slot_off.append(len(slot_var))
    # partition: first point index whose cumulative weight (measurements + 4 per point) reaches total * r / world
    w = off + 4 * np.arange(n_pt + 1)
    part = [0] + [int(np.searchsorted(w[:n_pt], w[n_pt] * r // world, side="left")) for r in range(1, world)] + [n_pt]
    cnt = np.diff(off)
    pt_order = np.concatenate([lo + np.argsort(-cnt[lo:hi], kind="stable") for lo, hi in zip(part[:-1], part[1:])]) if n_pt else np.zeros(0, int)
    keyed = []
    for q in range(off[part[rank]], off[part[rank + 1]]):
        vo = meas_b[q, 0]
        if vo < 0:
            continue
        keyed.append((vo * npv + vo, q))
        if meas_b[q, 2]:
            vs = pt_info[meas_b[q, 3], 2]
            keyed.append((min(vo, vs) * npv + max(vo, vs), q))
    keyed.sort()
    items = []
    i = 0
    while i < len(keyed):
        j = i
        while j < len(keyed) and keyed[j][0] == keyed[i][0]:
            j += 1
        for b in range(i, j, 128):
            items.append((keyed[i][0] // npv, keyed[i][0] % npv, b, min(b + 128, j)))
        i = j
    K = np.diff(np.asarray(slot_off))[part[rank]:part[rank + 1]]
    return dict(pose_var=pose_var, meas_orig=order, pt_meas_off=off, slot_var=np.asarray(slot_var), slot_pt=np.asarray(slot_pt),
                pt_slot_off=np.asarray(slot_off), meas_b=meas_b, pt_info=pt_info, part_pt=np.asarray(part), pt_order=pt_order,
                pb_idx=np.asarray([q for _, q in keyed]), pb_items=np.asarray(items).reshape(-1, 4), n_inc=int((K * (K + 1) // 2).sum()))


@pytest.mark.parametrize("shuffle", [True, False])
@pytest.mark.parametrize("cfg,seed,world", [("tiny", 0, 1), ("tiny", 1, 2), ("cfg1", 0, 1), ("cfg1", 2, 3)])
def test_layout_matches_restatement(cfg, seed, world, shuffle):
    prob = synth.make_ba_config(cfg, seed=seed)
    # make the case less regular: shuffle the measurement order (the two-level sort path; point-major input takes the
    # no-sort path), fix a few points, fix a second pose
    rng = np.random.default_rng(seed)
    prob = copy.copy(prob)
    perm = rng.permutation(prob.n_meas) if shuffle else np.arange(prob.n_meas)
    for k in ("meas_xy", "meas_chain", "meas_pt", "meas_noise", "meas_cam"):
        setattr(prob, k, np.ascontiguousarray(np.asarray(getattr(prob, k))[perm]))
    pf = np.array(prob.pt_fixed, copy=True)
    pf[rng.choice(prob.n_pt, 5, replace=False)] = 1
    prob.pt_fixed = pf
    for rank in range(world):
        got, _ = capi.ba_prepare(prob, rank=rank, world=world)
        ref = _restate(prob, rank, world)
        for k, v in ref.items():
            assert np.array_equal(np.asarray(got[k]), v), (k, rank)
        m = got["meas_orig"]
        assert np.array_equal(got["meas_xy"], np.asarray(prob.meas_xy)[m])
        assert np.array_equal(got["meas_info"], 1.0 / np.sqrt(np.asarray(prob.meas_noise)[m]))
        assert np.array_equal(got["meas_a"][:, 0], np.asarray(prob.meas_chain)[m, 0])
        assert np.array_equal(got["meas_a"][:, 2], np.asarray(prob.meas_cam)[m])
        assert np.array_equal(got["meas_a"][:, 3], m)
        assert got["max_slots"] == np.diff(got["pt_slot_off"]).max()
        assert np.array_equal(got["part_pt"], capi_partition(prob, world))


@pytest.mark.parametrize("shuffle", [True, False])
@pytest.mark.parametrize("threads", [2, 4, 7])
def test_threaded_passes_give_identical_layout(threads, shuffle):
    """the marshalling passes run on several host threads for large maps; the output does not depend on the count"""
    prob = synth.make_ba_config("cfg1", seed=1)
    rng = np.random.default_rng(5)
    prob = copy.copy(prob)
    perm = rng.permutation(prob.n_meas) if shuffle else np.arange(prob.n_meas)
    for k in ("meas_xy", "meas_chain", "meas_pt", "meas_noise", "meas_cam"):
        setattr(prob, k, np.ascontiguousarray(np.asarray(getattr(prob, k))[perm]))
    for world, rank in [(1, 0), (2, 1)]:
        one, _ = capi.ba_prepare(prob, rank=rank, world=world, want_rows=True, threads=1, par_min_meas=16384)
        many, _ = capi.ba_prepare(prob, rank=rank, world=world, want_rows=True, threads=threads, par_min_meas=0)
        for k, v in one.items():
            assert np.array_equal(np.asarray(many[k]), np.asarray(v)), k
    # the first offending measurement is reported whichever thread finds it
    bad = copy.copy(prob)
    a = np.array(prob.meas_cam, copy=True)
    a[[prob.n_meas - 2, prob.n_meas // 2 + 1, 17]] = 99
    bad.meas_cam = a
    with pytest.raises(capi.McpError) as e:
        capi.ba_prepare(bad, threads=threads, par_min_meas=0)
    assert "measurement 17:" in str(e.value)
    capi.ba_prepare(prob, threads=1, par_min_meas=16384)


def test_thread_pool_survives_many_back_to_back_loads():
    """mcp_ba_load runs once per BundleAdjust call for the lifetime of the process: thousands of pool dispatches, also
    with more threads than cores and after the pool has been re-created, must neither hang nor change the result."""
    prob = synth.make_ba_config("tiny", seed=0)
    one, _ = capi.ba_prepare(prob, threads=1, par_min_meas=1 << 30)
    for threads in (8, 3, 32, 2):
        many, _ = capi.ba_prepare(prob, threads=threads, par_min_meas=0, reps=1500)
        for k, v in one.items():
            assert np.array_equal(np.asarray(many[k]), np.asarray(v)), (k, threads)
    capi.ba_prepare(prob, threads=1, par_min_meas=16384)


def capi_partition(prob, world):
    """the same partition through a pure numpy statement of SURVEY.md §8(e): contiguous, measurement-count balanced"""
    off = np.concatenate([[0], np.cumsum(np.bincount(np.asarray(prob.meas_pt), minlength=prob.n_pt))])
    w = off + 4 * np.arange(prob.n_pt + 1)
    return np.asarray([0] + [int(np.searchsorted(w[:prob.n_pt], w[prob.n_pt] * r // world)) for r in range(1, world)] + [prob.n_pt])


def test_row_lists_cover_every_slot_once():
    """k_schur_rows work lists (MCP_BA_SCHUR=0): every (point, slot) entry of the rank appears exactly once, grouped by
    pose variable, groups fit the staging buffer and items tile the groups."""
    prob = synth.make_ba_config("cfg1", seed=0)
    for world, rank in [(1, 0), (2, 1)]:
        got, _ = capi.ba_prepare(prob, rank=rank, world=world, want_rows=True)
        s_lo, s_hi = got["pt_slot_off"][got["part_pt"][rank]], got["pt_slot_off"][got["part_pt"][rank + 1]]
        ent = got["rs_ent"]
        assert sorted(ent[:, 0].tolist()) == list(range(s_lo, s_hi))
        end_of_point = got["pt_slot_off"][got["slot_pt"][ent[:, 0]] + 1]
        assert np.array_equal(ent[:, 1], end_of_point - ent[:, 0])
        first, cnt = got["rs_grp"] >> 4, got["rs_grp"] & 15
        assert np.array_equal(first, np.concatenate([[0], np.cumsum(cnt)[:-1]])) and cnt.min() >= 1 and cnt.max() <= 8
        for f, c in zip(first, cnt):
            assert (192 + 144 * ent[f:f + c, 1]).sum() <= 6144 or c == 1
        it = got["rs_items"]
        assert it[0, 1] == 0 and it[-1, 2] == len(first) and np.array_equal(it[1:, 1], it[:-1, 2])
        for a, g0, g1, _ in it:
            e = ent[first[g0]:first[g1 - 1] + cnt[g1 - 1]]
            assert np.all(got["slot_var"][e[:, 0]] == a)
            assert np.all(got["slot_var"][e[:, 0] + e[:, 1] - 1] - a < got["rs_nblk"])


def test_argument_errors():
    prob = synth.make_ba_config("tiny", seed=0)
    for field, idx, val, what in [("meas_pt", 3, prob.n_pt, "point index"), ("meas_cam", 0, 99, "camera index"),
                                  ("meas_noise", 1, 0.0, "noise"), ("meas_chain", (2, 0), prob.n_pose, "chain index"),
                                  ("pt_chain", (0, 0), -1, "chain index")]:
        bad = copy.copy(prob)
        a = np.array(getattr(prob, field), copy=True)
        a[idx] = val
        setattr(bad, field, a)
        with pytest.raises(capi.McpError) as e:
            capi.ba_prepare(bad)
        assert what in str(e.value)
    # a movable second chain link (calibration BA) is rejected, as mcp_ba_load documents
    bad = copy.copy(prob)
    a = np.array(prob.meas_chain, copy=True)
    second = int(a[0, 1])
    if second >= 0:
        pf = np.array(prob.pose_fixed, copy=True)
        pf[second] = 0
        bad.pose_fixed = pf
        with pytest.raises(capi.McpError) as e:
            capi.ba_prepare(bad)
        assert "not supported" in str(e.value)


def test_empty_and_degenerate_maps():
    prob = synth.make_ba_config("tiny", seed=0)
    e = copy.copy(prob)
    for k in ("meas_xy", "meas_chain", "meas_pt", "meas_noise", "meas_cam"):
        setattr(e, k, np.asarray(getattr(prob, k))[:0])
    got, _ = capi.ba_prepare(e)
    assert got["n_slots"] == 0 and len(got["pb_idx"]) == 0 and len(got["pb_items"]) == 0 and got["n_inc"] == 0
    assert np.array_equal(got["pt_meas_off"], np.zeros(prob.n_pt + 1, int))
    # everything fixed: no slots, no pose blocks
    f = copy.copy(prob)
    f.pose_fixed = np.ones_like(np.asarray(prob.pose_fixed))
    got, _ = capi.ba_prepare(f)
    assert got["npv"] == 0 and got["n_slots"] == 0 and len(got["pb_idx"]) == 0
