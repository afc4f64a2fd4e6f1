"""K4 driver for profiling / tuning: one batched local-neighbourhood call on a regular grid (both variables co-located).

    python tools/k4_run.py [--nx 100] [--m 20000] [--md 0.08] [--reps 3] [--metric euclid|haversine]

Prints targets/s and FP64 TFLOP/s (k^3/3 + 2 k^2 per target).  Under ncu: add --reps 1 and filter -k regex:ck_local_predict.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sif-xco2-cokriging_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=100)
    ap.add_argument("--m", type=int, default=20000)
    ap.add_argument("--md", type=float, default=0.08)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--metric", default="euclid")
    ap.add_argument("--nu", type=float, default=1.5)
    ap.add_argument("--gather", action="store_true", help="gather the local matrices from a precomputed joint covariance")
    ap.add_argument("--phases", action="store_true", help="per-phase cycle shares from the kernel's debug counters")
    args = ap.parse_args()
    import torch
    from cokrig_b200 import METRIC_EUCLID, METRIC_HAVERSINE, ops
    nu = args.nu
    if args.metric == "euclid":
        gx = np.linspace(0, 1, args.nx)
        grid = np.array([(x, y) for y in gx for x in gx])
        pc = np.random.default_rng(7).uniform(0, 1, (args.m, 2))
        params, metric, md = [1, 1, nu, nu, nu, .2, .2, .2, .01, .01, -.6], METRIC_EUCLID, args.md
    else:  # the same geometry mapped to a 20 x 20 degree box: md in km
        gx = np.linspace(0, 20, args.nx)
        grid = np.array([(30 + y, -100 + x) for y in gx for x in gx])
        u = np.random.default_rng(7).uniform(0, 20, (args.m, 2))
        pc = np.c_[30 + u[:, 0], -100 + u[:, 1]]
        params, metric, md = [1, 1, nu, nu, nu, 400., 400., 400., .01, .01, -.6], METRIC_HAVERSINE, args.md * 2000.0
    z = [np.random.default_rng(k).standard_normal(len(grid)) for k in (1, 2)]
    cd = [ops.coords_to_device(grid), ops.coords_to_device(grid)]
    zd = [ops.to_device(v) for v in z]
    pd_ = ops.coords_to_device(pc)
    sigma = ops.joint_cov(cd, params, 2, metric) if args.gather else None
    dbg = None
    if args.phases:
        from cokrig_b200 import _lib
        dbg = torch.zeros(6, dtype=torch.int64, device="cuda")
        _lib.lib.ck_local_debug_buffer(dbg.data_ptr())
    best, k = 1e9, None
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, _, k, _ = ops.local_predict(cd, zd, pd_, params, 2, 1, metric, md, sigma=sigma)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    kf = k.astype(float)
    if dbg is not None:
        c = dbg.cpu().numpy().astype(float)
        names = ["scan", "cov_entries", "main_loop", "diag_block", "panel", "final"]
        print(json.dumps({"phase_share": {n: round(float(v / c.sum()), 4) for n, v in zip(names, c)},
                          "cycles_per_target_per_cta": float(c.sum() / args.m / args.reps)}))
    print(json.dumps({"targets": args.m, "ms": best, "targets_per_s": args.m / best * 1e3, "k_mean": float(kf.mean()),
                      "k_max": int(k.max()), "TFs": float(np.sum(kf ** 3 / 3 + 2 * kf ** 2)) / best / 1e9, "metric": args.metric,
                      "nu": nu, "gather": bool(args.gather)}))


if __name__ == "__main__":
    main()
