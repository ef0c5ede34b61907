"""Host side of BA problem creation (pmv_ba_index_observations, no GPU): the caller's observation list must come out in
device order -- sorted by (window, point, camera, original index) -- whichever of the three routes builds it: the list
taken in place (already ordered), the window-by-window counting sort (>= 64 windows, e.g. the camera-major order
CeresBundleAdjustment.cpp:27-52 produces), or the general scatter + per-point sorts.  Reference: numpy lexsort."""
import numpy as np
import pytest

import pmv_b200


def _make(W, Nc, Np, seed, dup=False):
    rng = np.random.default_rng(seed)
    obs, cam, pt, off = [], [], [], [0]
    for w in range(W):
        vis = rng.random((Np, Nc)) < 0.45
        vis[rng.integers(0, Np)] = False                     # an unobserved point
        p, c = np.nonzero(vis)                               # (point, camera) order
        if dup:                                              # the same camera seeing a point twice
            p = np.concatenate([p, p[:3]]); c = np.concatenate([c, c[:3]])
        o = rng.normal(size=(len(p), 2)) * 100
        obs.append(o); cam.append(c.astype(np.int32)); pt.append(p.astype(np.int32)); off.append(off[-1] + len(p))
    return obs, cam, pt, np.array(off, np.int32)


def _reorder(obs, cam, pt, how, seed):
    rng = np.random.default_rng(seed)
    out = ([], [], [])
    for o, c, p in zip(obs, cam, pt):
        if how == "point_major":
            idx = np.lexsort((c, p))
        elif how == "camera_major":
            idx = np.lexsort((p, c))
        else:
            idx = rng.permutation(len(c))
        out[0].append(o[idx]); out[1].append(c[idx]); out[2].append(p[idx])
    return np.concatenate(out[0]), np.concatenate(out[1]), np.concatenate(out[2])


def _expected(obs, cam, pt, off, Nc, Np):
    W = len(off) - 1
    win = np.repeat(np.arange(W), np.diff(off)).astype(np.int32)
    order = np.lexsort((np.arange(len(cam)), cam, pt, win))          # stable: original index last
    pt_off = np.zeros(W * Np + 1, np.int64)
    np.add.at(pt_off, win.astype(np.int64) * Np + pt + 1, 1)
    cam_off = np.zeros(W * Nc + 1, np.int64)
    np.add.at(cam_off, win.astype(np.int64) * Nc + cam + 1, 1)
    e = {"cam": cam[order], "pt": pt[order], "win": win[order], "obs": obs[order], "pt_off": np.cumsum(pt_off), "cam_off": np.cumsum(cam_off)}
    key = e["win"].astype(np.int64) * Nc + e["cam"]
    e["cam_obs"] = np.argsort(key, kind="stable")                      # device indices grouped by (window, camera), ascending
    return e


@pytest.mark.parametrize("W,how,route", [(1, "point_major", 0), (1, "camera_major", 2), (1, "shuffled", 2),
                                         (5, "point_major", 0), (5, "shuffled", 2),
                                         (70, "point_major", 0), (70, "camera_major", 1), (70, "shuffled", 1)])
def test_device_order_by_every_route(W, how, route):
    Nc, Np = 6, 40
    obs, cam, pt, off = _make(W, Nc, Np, 7 + W, dup=(how == "shuffled"))
    o, c, p = _reorder(obs, cam, pt, how, 3)
    got = pmv_b200.ba_index_observations(o, c, p, Nc, Np, off if W > 1 else None)
    want = _expected(o, c, p, off, Nc, Np)
    assert got["route"] == route
    for k in ("pt_off", "cam_off", "cam", "pt", "win", "cam_obs"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["obs"], want["obs"])


def test_bad_lists_are_rejected():
    Nc, Np = 4, 10
    obs, cam, pt, off = _make(3, Nc, Np, 1)
    o, c, p = _reorder(obs, cam, pt, "point_major", 0)
    bad = c.copy(); bad[5] = Nc
    with pytest.raises(ValueError):
        pmv_b200.ba_index_observations(o, bad, p, Nc, Np, off)
    bad = p.copy(); bad[0] = -1
    with pytest.raises(ValueError):
        pmv_b200.ba_index_observations(o, c, bad, Nc, Np, off)
    off2 = off.copy(); off2[0] = 1                                     # does not cover the list
    with pytest.raises(ValueError):
        pmv_b200.ba_index_observations(o, c, p, Nc, Np, off2)
    off3 = off.copy(); off3[1], off3[2] = off[2], off[1]               # decreasing
    with pytest.raises(ValueError):
        pmv_b200.ba_index_observations(o, c, p, Nc, Np, off3)
