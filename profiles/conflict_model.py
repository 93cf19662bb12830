#!/usr/bin/env python
"""CPU model of the shared-memory gathers of k_pair_sum: builds the work units, lane orders and per-lane candidate lists of
the bench lattice exactly as grid.cu / k_units / k_unit_prep / k_pair_lists do, then counts -- for every list trip -- the
wavefronts the LDS.128 gathers of the staged records need (quarter-warp rule: the 8 lanes of a quarter are served
together; distinct 16-byte addresses in the same bank group serialise; equal addresses broadcast).

Used to choose the lane order and the record stride without GPU time (round 2). Reports
  conflict = wavefronts / ideal wavefronts (1.0 = conflict-free; ncu round 1: 1707 M / 894 M = 1.91)
  lane_util = listed pairs / (32 x warp trips x W)
  barrier_loss = sum over chunks of (slowest warp - mean warp) / sum over chunks of slowest warp
Usage: python profiles/conflict_model.py [n_target] [jitter]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from opensph_b200 import workloads  # noqa: E402

TILE_T = 128
R = 2.0
ETA = 1.3
ZBINS = 32


def chunk_layers(k):
    return [(2 * k - 2, 2 * k + 3), (2 * k - 1, 2 * k + 2), (2 * k, 2 * k + 1)]


def build(n_target, jitter=0.0, seed=1):
    pos, hl = workloads.hexagonal_sphere(n_target, 5.0e4)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        pos = pos + rng.uniform(-jitter, jitter, pos.shape) * hl
    h = ETA * hl
    a = R * h * (1 + 1e-6)
    lo = pos.min(0) - 1e-9
    ci = np.floor((pos - lo) / np.array([a, a, 0.5 * a])).astype(int)
    dim = ci.max(0) + 1
    key = (ci[:, 2] * dim[1] + ci[:, 1]) * dim[0] + ci[:, 0]
    order = np.lexsort((np.arange(len(pos)), pos[:, 0], key))  # cell, then x, then slot (k_sort_cells)
    pos, ci, key = pos[order], ci[order], key[order]
    start = np.searchsorted(key, np.arange(int(dim.prod()) + 1))
    return pos, h, a, lo, dim, start


def units_of_row(start, dim, k, cy):
    rbL = ((2 * k) * dim[1] + cy) * dim[0]
    hasU = 2 * k + 1 < dim[2]
    rbU = ((2 * k + 1) * dim[1] + cy) * dim[0] if hasU else 0
    seq = []
    cols = []
    for c in range(dim[0]):
        lo_ = np.arange(start[rbL + c], start[rbL + c + 1])
        up_ = np.arange(start[rbU + c], start[rbU + c + 1]) if hasU else np.zeros(0, int)
        seq.append(np.concatenate([lo_, up_]))
        cols.append(np.full(len(lo_) + len(up_), c))
    seq = np.concatenate(seq).astype(int)
    cols = np.concatenate(cols).astype(int)
    out = []
    for b in range(0, len(seq), TILE_T):
        idx = seq[b:b + TILE_T]
        cc = cols[b:b + TILE_T]
        out.append((idx, int(cc.min()), int(cc.max())))
    return out


def lane_perm(P, a, lo, k, cy, mode):
    """Order of the unit's targets (indices into P) for lane assignment."""
    n = len(P)
    zrel = P[:, 2] - (lo[2] + 2 * k * 0.5 * a)
    zbin = np.clip((zrel * (ZBINS / a)).astype(int), 0, ZBINS - 1)
    if mode == "z":  # round 1: stable counting sort by z bin (ties keep the column order)
        return np.argsort(zbin, kind="stable")
    yrel = P[:, 1] - (lo[1] + cy * a)
    if mode.startswith("zyx") and mode != "zyxw":  # z bin, then y bin, then x
        nb = int(mode[3:]) if len(mode) > 3 else 16
        ybin = np.clip((yrel * (nb / a)).astype(int), 0, nb - 1)
        zb = np.clip((zrel * (nb / a)).astype(int), 0, nb - 1)
        return np.lexsort((P[:, 0], ybin, zb))
    if mode == "zx":  # z-bands of 32 (as "z"), lanes inside a warp by x
        pz = np.argsort(zbin, kind="stable")
        out = []
        for w in range(0, n, 32):
            s = pz[w:w + 32]
            out.append(s[np.argsort(P[s, 0], kind="stable")])
        return np.concatenate(out)
    if mode == "zyxw":  # z-bands of 32 (as "z"), lanes inside a warp by (y bin, x)
        pz = np.argsort(zbin, kind="stable")
        ybin = np.clip((yrel * (8 / a)).astype(int), 0, 7)
        out = []
        for w in range(0, n, 32):
            s = pz[w:w + 32]
            out.append(s[np.lexsort((P[s, 0], ybin[s]))])
        return np.concatenate(out)
    raise ValueError(mode)


def model(n_target, mode="z", W=3, stride_chunks=9, jitter=0.0, max_units=400, list_sort=None):
    pos, h, a, lo, dim, start = build(n_target, jitter)
    reach2 = (R * h) ** 2
    rng = np.random.default_rng(7)
    rows = [(k, cy) for k in range((dim[2] + 1) // 2) for cy in range(dim[1])]
    rng.shuffle(rows)
    tot = dict(wave=0, ideal=0, pairs=0, trips=0, units=0, slow=0.0, mean=0.0)
    cls = {c: [0, 0] for c in range(3)}
    for (k, cy) in rows:
        for (idx, cA, cB) in units_of_row(start, dim, k, cy):
            if len(idx) < TILE_T:
                continue
            P = pos[idx]
            perm = lane_perm(P, a, lo, k, cy, mode)
            P, tix = P[perm], idx[perm]
            x0, x1 = max(cA - 1, 0), min(cB + 1, dim[0] - 1)
            for chn, (zl, zh) in enumerate(chunk_layers(k)):
                # staged records of the chunk: six rows, contiguous
                staged = []
                for z in (zl, zh):
                    for dy in (-1, 0, 1):
                        y = cy + dy
                        if z < 0 or z >= dim[2] or y < 0 or y >= dim[1]:
                            continue
                        base = (z * dim[1] + y) * dim[0]
                        staged.append(np.arange(start[base + x0], start[base + x1 + 1]))
                staged = np.concatenate(staged) if staged else np.zeros(0, int)
                if len(staged) == 0:
                    continue
                C = pos[staged]
                d2 = ((P[:, None, :] - C[None, :, :]) ** 2).sum(-1)
                hit = d2 < reach2 * (1 + 4e-5)  # FP32 filter ~ exact here (uniform h); includes the target itself
                cnt = hit.sum(1)
                L = np.full((TILE_T, int(cnt.max()) + W), -1, int)
                if list_sort == "mod8coop":
                    coop_schedule(hit, L)
                for t in range(TILE_T if list_sort != "mod8coop" else 0):
                    e = np.nonzero(hit[t])[0]
                    if list_sort == "mod8":  # order the lane's entries so that entry q has residue (q + lane) mod 8 when possible
                        e = reorder_mod8(e, t)
                    elif list_sort == "mod8skip":
                        e = reorder_mod8skip(e, t)
                    elif list_sort == "mod8last":
                        e = reorder_mod8(e, t, True)
                    L[t, :len(e)] = e
                warp_cost = np.zeros(4)
                for w in range(4):
                    lw = L[32 * w:32 * w + 32]
                    cw = cnt[32 * w:32 * w + 32]
                    ntrip = int(np.ceil(cw.max() / W)) if cw.max() > 0 else 0
                    warp_cost[w] = ntrip
                    tot["trips"] += ntrip
                    tot["pairs"] += int(cw.sum())
                    for tr in range(ntrip):
                        for e in range(W):
                            q = tr * W + e
                            rec = lw[:, q]
                            # a lane takes part in the trip while it has entries left (W-wide loop, then the tails)
                            active = (tr * W) < cw
                            act = active & (rec >= 0)
                            # lanes inside the W-wide loop whose entry q is beyond the list do not load (tail handled 2 + 1)
                            for qw in range(4):
                                sl = slice(8 * qw, 8 * qw + 8)
                                r = rec[sl][act[sl]]
                                if len(r) == 0:
                                    continue
                                ur = np.unique(r)
                                # the nine 16-byte pieces of a record: piece c of record r sits in bank group (stride r + c) mod 8
                                grp = (stride_chunks * ur) % 8
                                mult = np.bincount(grp, minlength=8).max()
                                tot["wave"] += 9 * mult
                                tot["ideal"] += 9
                                cls[chn][0] += mult
                                cls[chn][1] += 1
                tot["slow"] += warp_cost.max()
                tot["mean"] += warp_cost.mean()
            tot["units"] += 1
            if tot["units"] >= max_units:
                break
        if tot["units"] >= max_units:
            break
    return {
        "mode": mode, "W": W, "stride": stride_chunks, "jitter": jitter, "list_sort": list_sort, "units": tot["units"],
        "conflict": tot["wave"] / max(tot["ideal"], 1),
        "lane_util": tot["pairs"] / max(32 * tot["trips"] * W, 1),
        "barrier_loss": (tot["slow"] - tot["mean"]) / max(tot["slow"], 1),
        "pairs_per_target": tot["pairs"] / max(tot["units"] * TILE_T, 1),
        "by_chunk": [round(cls[c][0] / max(cls[c][1], 1), 3) for c in range(3)], "share": [round(cls[c][1] / max(sum(v[1] for v in cls.values()), 1), 3) for c in range(3)],
    }


def coop_schedule(hit, L):
    """Quarter-cooperative variant: at slot q lane l prefers residue (q + l) mod 8; a lane whose preferred bucket is
    empty takes its fullest bucket among the residues no honouring lane of its quarter uses at this slot (ballot)."""
    for qw in range(TILE_T // 8):
        lanes = range(8 * qw, 8 * qw + 8)
        buckets = {t: [list(np.nonzero(hit[t])[0][(np.nonzero(hit[t])[0] % 8) == b]) for b in range(8)] for t in lanes}
        left = {t: sum(len(b) for b in buckets[t]) for t in lanes}
        q = 0
        while any(left[t] > 0 for t in lanes):
            taken = set()
            fb = []
            for t in lanes:
                if left[t] == 0:
                    continue
                b = (q + t) % 8
                if buckets[t][b]:
                    L[t, q] = buckets[t][b].pop(0)
                    left[t] -= 1
                    taken.add(b)
                else:
                    fb.append(t)
            for t in fb:
                free = [b for b in range(8) if b not in taken and buckets[t][b]]
                cand = free if free else [b for b in range(8) if buckets[t][b]]
                bb = max(cand, key=lambda i: len(buckets[t][i]))
                L[t, q] = buckets[t][bb].pop(0)
                left[t] -= 1
                taken.add(bb)
            q += 1


def reorder_mod8skip(e, lane):
    """Entries sorted by (position inside the residue bucket, (residue - lane) mod 8): the Latin-square slots with the
    holes squeezed out."""
    b = e % 8
    pos = np.zeros(len(e), int)
    cnt = np.zeros(8, int)
    for i, bb in enumerate(b):
        pos[i] = cnt[bb]
        cnt[bb] += 1
    key = pos * 8 + ((b - lane) % 8)
    return e[np.argsort(key, kind="stable")]


def reorder_mod8(e, lane, last=False):
    buckets = [list(e[(e % 8) == b]) for b in range(8)]
    out = []
    q = 0
    left = len(e)
    while left > 0:
        b = (q + lane) % 8
        if buckets[b]:
            out.append(buckets[b].pop(0))
        else:  # take from the fullest bucket
            bb = max(range(8), key=lambda i: len(buckets[i]))
            out.append(buckets[bb].pop(0))
        left -= 1
        q += 1
    return np.array(out, int)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
    jit = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    for mode in ("z", "zx", "zyxw", "zyx8", "zyx16"):
        for ls in (None, "mod8"):
            r = model(n, mode, 3, 9, jit, 150, ls)
            print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
