"""Parity measures shared by tests/ and bench.py (harness code, not product).

north_star: "the same corner set as the reference extractor up to response ties within 1e-5".  Two corner lists can
legitimately differ where two responses are closer than the arithmetic difference between two correct
implementations (ours sums exact integers; OpenCV sums rounded fp32 products): the ORDER of two near-tied corners
swaps, and -- through the greedy min-distance pass -- a swap can change which of two neighbours survives.  The checks
below turn that sentence into numbers instead of a boolean."""
from __future__ import annotations

import numpy as np

TIE = 1e-5


def list_parity(got_xy, want_xy, eig, tie=TIE):
    """Compare two ordered corner lists given the reference response map `eig` (cv2.cornerMinEigenVal).
    Returns: identical, set_equal, symmetric_difference, first_differing_rank and max_swapped_gap_rel = the largest
    response difference (in units of max(eig)) between any two corners whose relative order differs in the two lists."""
    g = np.asarray(got_xy).astype(np.int64).reshape(-1, 2); w = np.asarray(want_xy).astype(np.int64).reshape(-1, 2)
    vmax = float(eig.max())
    out = {"n_got": int(len(g)), "n_want": int(len(w)), "identical": bool(g.shape == w.shape and np.array_equal(g, w))}
    kg = g[:, 1] * eig.shape[1] + g[:, 0]; kw = w[:, 1] * eig.shape[1] + w[:, 0]
    sg, sw = set(kg.tolist()), set(kw.tolist())
    out["set_equal"] = sg == sw
    out["symmetric_difference"] = len(sg ^ sw)
    nd = np.nonzero(kg[:min(len(kg), len(kw))] != kw[:min(len(kg), len(kw))])[0]
    out["first_differing_rank"] = int(nd[0]) if len(nd) else (None if len(kg) == len(kw) else min(len(kg), len(kw)))
    # inversions among the common corners: A = want order (sorted by eig desc), perm = rank in got
    common = [k for k in kw.tolist() if k in sg]
    rank_g = {k: i for i, k in enumerate(kg.tolist())}
    perm = np.array([rank_g[k] for k in common], np.int64)
    ev = np.array([eig.flat[k] for k in common], np.float64)
    gap = 0.0
    if len(perm) > 1:
        order = np.argsort(perm, kind="stable")           # positions (in want order) sorted by got rank
        # for position i: the furthest later position j > i with perm[j] < perm[i]
        pm = np.maximum.accumulate(order)                 # prefix max of want-positions over got ranks 0..r
        r_of = np.empty_like(order); r_of[order] = np.arange(len(order))
        for i in range(len(perm)):
            r = r_of[i]
            if r > 0:
                j = pm[r - 1]
                if j > i:
                    gap = max(gap, abs(ev[i] - ev[j]))
    out["max_swapped_gap_rel"] = gap / vmax if vmax > 0 else 0.0
    out["swaps_within_tie"] = bool(out["max_swapped_gap_rel"] <= tie)
    return out


def gftt_valid_up_to_ties(img, xy, max_corners, quality, min_dist, tie=TIE, eig=None):
    """Is `xy` a valid output of cv::goodFeaturesToTrack(img, max_corners, quality, min_dist) if responses closer than
    tie * max may compare either way?  (i) every corner is a candidate (above threshold, 3x3 local maximum, not on the
    rim), (ii) the list is sorted by response, (iii) all pairs keep min_dist, (iv) every stronger candidate that is
    missing lies within min_dist of an accepted corner that is at least as strong -- each up to the tolerance.
    Returns (ok, reasons)."""
    import cv2
    from scipy.spatial import cKDTree
    if eig is None:
        eig = cv2.cornerMinEigenVal(img, 3, ksize=3)
    H, W = eig.shape
    vmax = float(eig.max()); tau = tie * vmax; thr = vmax * quality
    dil = cv2.dilate(eig, np.ones((3, 3), np.uint8))
    xy = np.asarray(xy).astype(np.int64).reshape(-1, 2)
    reasons = []
    e = eig[xy[:, 1], xy[:, 0]].astype(np.float64)
    inside = (xy[:, 0] >= 1) & (xy[:, 0] <= W - 2) & (xy[:, 1] >= 1) & (xy[:, 1] <= H - 2)
    if not inside.all():
        reasons.append(f"{int((~inside).sum())} corners on the rim")
    bad = (e < thr - tau) | (e < dil[xy[:, 1], xy[:, 0]] - tau)
    if bad.any():
        reasons.append(f"{int(bad.sum())} corners are not candidates")
    if len(e) > 1 and (np.diff(e) > tau).any():
        reasons.append(f"{int((np.diff(e) > tau).sum())} order violations beyond the tie tolerance")
    if min_dist >= 1 and len(xy) > 1:
        t = cKDTree(xy)
        pairs = t.query_pairs(float(min_dist) - 1e-9)
        close = [(a, b) for a, b in pairs if (xy[a] - xy[b]) @ (xy[a] - xy[b]) < min_dist * min_dist]
        if close:
            reasons.append(f"{len(close)} pairs closer than min_dist")
    # maximality
    cand = (eig > thr + tau) & (eig >= dil)
    cand[0, :] = cand[-1, :] = False; cand[:, 0] = cand[:, -1] = False
    cy, cx = np.nonzero(cand)
    have = set((xy[:, 1] * W + xy[:, 0]).tolist())
    floor = e.min() if (max_corners > 0 and len(xy) >= max_corners) else -np.inf   # truncated list: only stronger ones matter
    miss = [(x, y) for x, y in zip(cx.tolist(), cy.tolist()) if (y * W + x) not in have and eig[y, x] > floor + tau]
    if miss:
        if min_dist < 1 or len(xy) == 0:
            reasons.append(f"{len(miss)} stronger candidates missing")
        else:
            t = cKDTree(xy)
            unexplained = 0
            for (x, y) in miss:
                idx = t.query_ball_point([x, y], float(min_dist))
                ok = any(((xy[i, 0] - x) ** 2 + (xy[i, 1] - y) ** 2) < min_dist * min_dist and e[i] >= eig[y, x] - tau for i in idx)
                unexplained += 0 if ok else 1
            if unexplained:
                reasons.append(f"{unexplained} stronger candidates missing without an accepted neighbour")
    return (not reasons), reasons
