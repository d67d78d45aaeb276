#!/usr/bin/env python
"""Phase timeline of one CG iteration from in-kernel %globaltimer stamps (sofab200_ctx_trace_begin/end): where the time
of the element pass and of the fused CG tail goes, per CTA.  Diagnostics used while tuning; not a bench line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sofa_b200 as sb  # noqa: E402
from bench import SCENE, build_mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--out", default="")
args = ap.parse_args()
dtype, template = (np.float32, "B200Vec3f") if args.dtype == "f32" else (np.float64, "B200Vec3d")
pos, tets, fixed = build_mesh(args.workload)
ctx = sb.Context(0)
mo = sb.MechanicalObject(ctx, template, position=pos)
ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=SCENE["young"], poissonRatio=SCENE["poisson"], method="large", tileElems=args.tile)
mass = sb.DiagonalMass(mo, tets, massDensity=SCENE["density"])
node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=0.01, gravity=SCENE["gravity"], rayleighStiffness=0.1, rayleighMass=0.1,
                     iterations=25, tolerance=1e-9, threshold=1e-9)
for _ in range(4):
    node.step()
torch.cuda.synchronize()
ctx.trace_begin()
node.step()
tile, tail = ctx.trace_end()


def summarize(rec, names):
    rec = rec[rec[:, 0] > 0]
    if rec.shape[0] == 0:
        return None
    t0 = rec[:, 0].min()
    rel = (rec[:, :len(names)].astype(np.int64) - np.int64(t0)) / 1000.0   # us
    out = {"ctas": int(rec.shape[0]), "sms": int(len(np.unique(rec[:, 7])))}
    for i, nme in enumerate(names):
        c = rel[:, i]
        out[nme] = {"min": round(float(c.min()), 2), "median": round(float(np.median(c)), 2), "max": round(float(c.max()), 2)}
    d = np.diff(rel, axis=1)
    out["durations_median_us"] = {f"{names[i]}->{names[i + 1]}": round(float(np.median(d[:, i])), 2) for i in range(len(names) - 1)}
    return out, rel, rec[:, 7]


res = {}
s = summarize(tile, ["start", "phase1_done", "loop_done", "barrier_done", "end"])
if s:
    res["element_pass"] = s[0]
    rel, sm = s[1], s[2]
    # second CTA on the same SM (wave 2) starts when the first ends
    order = np.argsort(rel[:, 0])
    res["element_pass"]["first_wave_end_median"] = round(float(np.median(rel[order[:len(order) // 2], 4])), 2)
fused = os.environ.get("SOFAB200_CG_FUSED", "1") != "0"
if fused:
    # second-generation kernel (cg_fused.cuh): marks of the 10th iteration of the solve (mark 12: start of the 11th)
    tl = tail[tail[:, 0] > 0].astype(np.int64)
    us = lambda a, b: round(float(np.median(tl[:, b] - tl[:, a])) / 1000.0, 2)
    mx = lambda a, b: round(float(np.max(tl[:, b] - tl[:, a])) / 1000.0, 2)
    t0 = tl[:, 0].min()
    names = {0: "iteration start", 1: "tile 1 elements done", 2: "tile 1 interior sums done", 3: "all tiles done", 4: "S1 seen (element side)",
             5: "all units done (CTA barrier)", 10: "S2: CTA sums ready", 14: "S2 arrival issued (GT>0: S1 seen by the dedicated warps)",
             15: "S2 wait + sum done (GT>0: dedicated warps' units done)", 6: "S2 done", 13: "update done", 12: "next iteration start"}
    dist = {"ctas": int(tl.shape[0])}
    for i, nme in names.items():
        c = tl[:, i]
        if not np.all(c > 0):
            continue
        c = (c - t0) / 1000.0
        dist[nme] = {"min": round(float(c.min()), 2), "p10": round(float(np.percentile(c, 10)), 2), "median": round(float(np.median(c)), 2),
                     "p90": round(float(np.percentile(c, 90)), 2), "max": round(float(c.max()), 2), "argmax_cta": int(np.argmax(c))}
    res["cg_fused_iteration_10_marks_us"] = dist
    res["cg_fused_iteration_10_durations_us"] = {
        "first tile: elements": us(0, 1), "first tile: interior sums": us(1, 2), "remaining tiles (elements + interior sums)": us(2, 3),
        "all tiles done -> S1 seen": us(3, 4), "S1 seen -> all units done (CTA barrier)": us(4, 5), "S2 (post, wait, sum)": us(5, 6),
        "S2 done -> update done": us(6, 13), "whole iteration (start -> next start), median / max CTA": [us(0, 12), mx(0, 12)]}
    res["cg_fused_kernel_us"] = {"tables": us(8, 9), "kernel start -> end (median CTA)": us(8, 11)}
    s = None
else:
    s = summarize(tail, ["iter_start", "tiles_done", "barrier1", "shared_done", "barrier2", "xr_done", "barrier3"])
if s:
    res["cg_persistent_last_iteration"] = s[0]      # the persistent CG kernel (the multi-kernel tail uses other marks)
    tl = tail[tail[:, 0] > 0].astype(np.int64)
    us = lambda a, b: round(float(np.median(tl[:, b] - tl[:, a])) / 1000.0, 2)
    res["cg_persistent_last_iteration"]["detail_us"] = {
        "p-update + staging (all tiles of the CTA)": us(0, 12), "first tile: elements": us(12, 13), "first tile: interior sums": us(13, 14)}
    res["cg_persistent_kernel"] = {"node tables -> shared memory": us(8, 9), "|b| and first rho (two grid syncs)": us(9, 10),
                                   "kernel start -> kernel end (median CTA)": us(8, 11), "x written back after the last sync": us(6, 11)}
print(json.dumps(res, indent=1))
if args.out:
    np.savez_compressed(args.out, tile=tile, tail=tail)
