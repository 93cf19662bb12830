#!/usr/bin/env python
"""CPU work model of the tiled pair kernel: counts, for the bench lattice, what every warp of every work unit would
execute under different lane assignments / phase-1 strategies, so that kernel designs can be compared without GPU time.

It mirrors grid.cu (cells a x a x a/2), k_units (<= 128 targets per double row) and the chunking of pair_tiled.cu
(far / near / centre layer pairs), then reports per warp-unit:
  p1   = phase-1 candidate tests issued at warp level (SIMT: the longest lane window of every candidate row counts)
  p2   = phase-2 trips (two pairs per trip; the longest lane list of every chunk counts)
  bar  = what the CTA-wide barrier per chunk costs: sum over chunks of (slowest warp - mean warp)
Usage: python profiles/work_model.py [n_target]
"""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
from opensph_b200 import workloads  # noqa: E402

TILE_T = 128
R = 2.0
ETA = 1.3
C_TEST_OLD, C_TEST_NEW, C_ROW, C_SEARCH, C_TRIP = 14, 10, 40, 90, 330  # warp instructions (phase-2 trip: 2 pairs)


def chunk_layers(k):
    return [(2 * k - 2, 2 * k + 3), (2 * k - 1, 2 * k + 2), (2 * k, 2 * k + 1)]


def model(n_target, lane_mode="z4", p1_mode="cells", pack="nextfit", tile_c=576):
    pos, hl = workloads.hexagonal_sphere(n_target, 5.0e4)
    h = ETA * hl
    a = R * h * (1 + 1e-6)
    lo = pos.min(0) - 1e-9
    ci = np.floor((pos - lo) / np.array([a, a, 0.5 * a])).astype(int)
    dim = ci.max(0) + 1
    key = (ci[:, 2] * dim[1] + ci[:, 1]) * dim[0] + ci[:, 0]
    order = np.lexsort((pos[:, 0], key)) if p1_mode != "cells" else np.argsort(key, kind="stable")
    pos, ci, key = pos[order], ci[order], key[order]
    ncell = int(dim.prod())
    start = np.searchsorted(key, np.arange(ncell + 1))
    reach = R * h
    tot = dict(p1=0.0, p2=0.0, bar=0.0, units=0, lanes_p1=0.0, lanes_p2=0.0, pipe=0.0, cta=0.0, tests=0.0)

    def row_range(z, y, x0, x1):
        if z < 0 or z >= dim[2] or y < 0 or y >= dim[1]:
            return 0, 0
        base = (z * dim[1] + y) * dim[0]
        return start[base + x0], start[base + x1 + 1]

    for k in range((dim[2] + 1) // 2):
        for cy in range(dim[1]):
            rbL = ((2 * k) * dim[1] + cy) * dim[0]
            hasU = 2 * k + 1 < dim[2]
            rbU = ((2 * k + 1) * dim[1] + cy) * dim[0] if hasU else 0
            colL = np.array([start[rbL + c + 1] - start[rbL + c] for c in range(dim[0])])
            colU = np.array([(start[rbU + c + 1] - start[rbU + c]) if hasU else 0 for c in range(dim[0])])
            col = colL + colU
            units = []  # (cA, cB, skip in the column-ordered sequence of cA.., number of targets)
            if pack == "nextfit":  # whole columns, close the unit when the next one does not fit (k_units, round 1)
                acc, cA, cLast = 0, 0, 0
                for c in range(dim[0]):
                    cnt = col[c]
                    if cnt == 0:
                        continue
                    if acc > 0 and (acc + cnt > TILE_T or c - cA + 3 > 20):
                        units.append((cA, cLast, 0, acc))
                        acc = 0
                    if acc == 0:
                        cA = c
                    acc += cnt
                    cLast = c
                    if acc >= TILE_T:
                        units.append((cA, cLast, 0, min(acc, TILE_T)))
                        acc = 0
                if acc > 0:
                    units.append((cA, cLast, 0, acc))
            else:  # columns may be split; "hybrid:T" closes at a column boundary once T targets are on board
                thr = int(pack.split(":")[1]) if ":" in pack else TILE_T
                c, skip = 0, 0
                while c < dim[0]:
                    while c < dim[0] and col[c] - skip == 0:
                        c, skip = c + 1, 0
                    if c >= dim[0]:
                        break
                    cA, skipA, taken, cLast = c, skip, 0, c
                    while c < dim[0] and taken < TILE_T and c - cA + 3 <= 20:
                        avail = col[c] - skip
                        if avail == 0:
                            c, skip = c + 1, 0
                            continue
                        if taken >= thr and skip == 0:
                            break
                        take = min(avail, TILE_T - taken)
                        taken += take
                        cLast = c
                        if take == avail:
                            c, skip = c + 1, 0
                        else:
                            skip += take
                            break
                    units.append((cA, cLast, skipA, taken))
            for (cA, cB, skipA, ntar) in units:
                seq = np.concatenate([np.concatenate([np.arange(start[rbL + c], start[rbL + c + 1]),
                                                      np.arange(start[rbU + c], start[rbU + c + 1]) if hasU else np.zeros(0, int)])
                                      for c in range(cA, cB + 1)]).astype(int)
                idx = seq[skipA:skipA + ntar]
                if len(idx) < 64:  # skip the sparse rim: the interior dominates at 10 M
                    continue
                n_live = len(idx)
                P = np.concatenate([pos[idx], np.full((TILE_T - n_live, 3), 1.0e9)])  # idle lanes: out of reach of everything
                idx = np.concatenate([idx, np.zeros(TILE_T - n_live, int)])
                tot["live"] = tot.get("live", 0) + n_live
                if lane_mode == "z4":
                    perm = np.argsort(P[:, 2], kind="stable")
                elif lane_mode == "z2x2":  # two z-bands (halves in z), each split in two x-halves
                    pz = np.argsort(P[:, 2], kind="stable")
                    perm = np.concatenate([h2[np.argsort(P[h2, 0], kind="stable")] for h2 in (pz[:64], pz[64:])])
                elif lane_mode == "zrr":  # z-ranks dealt round-robin to the four warps
                    pz = np.argsort(P[:, 2], kind="stable")
                    perm = np.concatenate([pz[w::4] for w in range(4)])
                elif lane_mode == "z2rr":  # two z-bands, each dealt round-robin to two warps
                    pz = np.argsort(P[:, 2], kind="stable")
                    perm = np.concatenate([pz[:64][0::2], pz[:64][1::2], pz[64:][0::2], pz[64:][1::2]])
                elif lane_mode == "z8":  # eight z-bands; warp w takes bands w and 7 - w (mirror images)
                    pz = np.argsort(P[:, 2], kind="stable").reshape(8, 16)
                    perm = np.concatenate([np.concatenate([pz[w], pz[7 - w]]) for w in range(4)])
                elif lane_mode == "x4":
                    perm = np.argsort(P[:, 0], kind="stable")
                elif lane_mode == "z4x":  # z-bands, lanes inside a warp ordered by x (same warps as z4)
                    pz = np.argsort(P[:, 2], kind="stable")
                    perm = np.concatenate([w[np.argsort(P[w, 0], kind="stable")] for w in pz.reshape(4, 32)])
                else:
                    raise ValueError(lane_mode)
                P = P[perm]
                cxl = ci[idx][perm][:, 0]
                x0, x1 = max(cA - 1, 0), min(cB + 1, dim[0] - 1)
                warp_chunk = np.zeros((4, 3))
                for (zl, zh) in chunk_layers(k):
                    tot_rows = sum(max(row_range(z, cy + dy, x0, x1)[1] - row_range(z, cy + dy, x0, x1)[0], 0) for z in (zl, zh) for dy in (-1, 0, 1))
                    tot["chunks"] = tot.get("chunks", 0) + max(1, math.ceil(tot_rows / tile_c))
                    tot["staged"] = tot.get("staged", 0) + tot_rows
                for ch, (zl, zh) in enumerate(chunk_layers(k)):
                    cnt_lane = np.zeros(TILE_T)
                    p1_warp = np.zeros(4)
                    for z in (zl, zh):
                        for dy in (-1, 0, 1):
                            y = cy + dy
                            b, e = row_range(z, y, x0, x1)
                            if e <= b:
                                continue
                            C = pos[b:e]
                            ylo, zlo = lo[1] + y * a, lo[2] + z * 0.5 * a
                            dyMin = np.maximum(np.maximum(ylo - P[:, 1], P[:, 1] - (ylo + a)), 0)
                            dzMin = np.maximum(np.maximum(zlo - P[:, 2], P[:, 2] - (zlo + 0.5 * a)), 0)
                            rem = reach * reach - dyMin ** 2 - dzMin ** 2
                            ext = np.sqrt(np.maximum(rem, 0))
                            d2 = ((P[:, None, :] - C[None, :, :]) ** 2).sum(-1)
                            cnt_lane += (d2 <= reach * reach).sum(1)
                            if p1_mode == "cells":
                                c0 = np.clip(np.floor((P[:, 0] - ext - lo[0]) / a), -1, dim[0]).astype(int)
                                c1 = np.clip(np.floor((P[:, 0] + ext - lo[0]) / a), -1, dim[0]).astype(int)
                                c0 = np.maximum(np.maximum(c0, cxl - 1), x0)
                                c1 = np.minimum(np.minimum(c1, cxl + 1), x1)
                                base = (z * dim[1] + y) * dim[0]
                                ln = np.where((rem > 0) & (c0 <= c1), start[base + np.clip(c1, 0, dim[0] - 1) + 1] - start[base + np.clip(c0, 0, dim[0] - 1)], 0)
                                ln = np.maximum(ln, 0)
                                ctest, crow = C_TEST_OLD, C_ROW
                            elif p1_mode.startswith("fine"):  # x-windows at the granularity of cells a / XSUB wide
                                xs = a / int(p1_mode[4:])
                                f0 = np.floor((P[:, 0] - ext - lo[0]) / xs)
                                f1 = np.floor((P[:, 0] + ext - lo[0]) / xs)
                                cf = np.floor((C[:, 0] - lo[0]) / xs)
                                ln = np.where(rem > 0, ((cf[None, :] >= f0[:, None]) & (cf[None, :] <= f1[:, None])).sum(1), 0)
                                ctest, crow = C_TEST_NEW, C_ROW
                            elif p1_mode == "exact":
                                ln = np.where(rem > 0, (np.abs(P[:, None, 0] - C[None, :, 0]) <= ext[:, None]).sum(1), 0)
                                ctest, crow = C_TEST_NEW, C_ROW + C_SEARCH
                            elif p1_mode == "bcast":  # union window of the warp, every lane tests every candidate in it
                                ln = np.zeros(TILE_T)
                                for w in range(4):
                                    s = slice(32 * w, 32 * w + 32)
                                    act = rem[s] > 0
                                    if act.any():
                                        xl = (P[s, 0] - ext[s])[act].min()
                                        xh = (P[s, 0] + ext[s])[act].max()
                                        ln[s] = ((C[:, 0] >= xl) & (C[:, 0] <= xh)).sum()
                                ctest, crow = C_TEST_NEW - 1, C_ROW
                            tot["tests"] += ln.sum()
                            lw = ln.reshape(4, 32)
                            mx = lw.max(1)
                            # SIMT cost of the 8-wide / 4-wide / tail loops
                            t8 = (lw // 8).max(1) * 8
                            t4 = ((lw % 8) // 4).max(1) * 4
                            t1 = (lw % 4).max(1)
                            p1_warp += (t8 + t4 + t1) * ctest + np.where(mx > 0, crow, 10)
                            tot["lanes_p1"] += lw.sum()
                            tot["p1"] += (t8 + t4 + t1).sum()
                    trips = np.ceil(cnt_lane.reshape(4, 32) / 2).max(1)
                    tot["p2"] += trips.sum()
                    tot["lanes_p2"] += cnt_lane.sum()
                    warp_chunk[:, ch] = p1_warp + trips * C_TRIP
                tot["bar"] += (warp_chunk.max(0) - warp_chunk.mean(0)).sum()
                tot["cta"] += warp_chunk.max(0).sum()
                tot["pipe"] += warp_chunk.sum(1).max()
                tot["units"] += 1
    u = tot["units"]
    wu = 4 * u
    return {
        "lane_mode": lane_mode, "p1_mode": p1_mode, "pack": pack, "units": u, "chunks_per_unit": tot["chunks"] / u,
        "staged_per_target": tot["staged"] / tot["live"], "instr_per_target": (tot["cta"] + 500 * tot["chunks"]) / tot["live"], "targets_per_unit": tot["live"] / u,
        "p1_tests_per_warp": tot["p1"] / wu, "p1_lane_util": tot["lanes_p1"] / (32 * tot["p1"]),
        "tests_per_lane": tot["tests"] / (u * TILE_T),
        "p2_trips_per_warp": tot["p2"] / wu, "p2_lane_util": tot["lanes_p2"] / (64 * tot["p2"]),
        "cta_instr_per_unit": tot["cta"] / u, "barrier_loss_frac": tot["bar"] / tot["cta"],
        "pipelined_instr_per_unit": tot["pipe"] / u,
    }


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
    for lane_mode, p1_mode in (("z4", "cells"), ("z4", "exact"), ("z4", "bcast"), ("z2x2", "exact"), ("z2x2", "bcast"), ("x4", "bcast"),
                               ("x4", "exact")):
        r = model(n, lane_mode, p1_mode)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
