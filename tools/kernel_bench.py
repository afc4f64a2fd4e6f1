"""Per-kernel timings on the BASELINE config shapes -- run on the GPU box.

    python tools/kernel_bench.py [--only k1,k2,k3,k4,nll] [--out gpurun_out/kernels.json] [--quick]

  k1  covariance assembly at C3 (N = 40 000, haversine) for half-integer and generic nu, and Euclidean
  k2  empirical variograms at C2 (10 000 cells / variable of the 0.05 deg CONUS lattice, 50 bins)
  k3  Cholesky / multi-RHS solve at several N (TFLOP/s, N^3/3 and N^2 m)
  k4  local-neighbourhood batch at C1 (40 x 40 grid, max_dist 0.2, 500 targets) and a larger batch
  nll Gaussian NLL objective evaluation at C3 (assemble + potrf + 1-RHS solve + logdet)
--quick shrinks every size (the ncu target: few, short launches).  All timings: CUDA events on the
current stream, best of 3 after one warm-up.  No oracle here: parity lives in tests/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "sif-xco2-cokriging_b200", "src"), os.path.join(ROOT, "sif-xco2-cokriging_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

HALF = [1.0, 0.8, 1.5, 1.5, 1.5, 500.0, 500.0, 500.0, 0.02, 0.02, -0.2]
GENERIC = [1.0, 0.8, 0.75, 1.0, 1.25, 500.0, 500.0, 500.0, 0.02, 0.02, -0.2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="k1,k2,k3,k4,nll")
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--k3-sizes", default="", help="comma-separated N for the k3 leg")
    args = ap.parse_args()
    only = set(args.only.split(","))

    import torch
    import fields
    from bench import make_workload
    from cokrig_b200 import METRIC_EUCLID, METRIC_HAVERSINE, _lib, ops
    torch.cuda.set_device(0)
    res = {"gpu": torch.cuda.get_device_name(0), "quick": args.quick}

    def timed(fn, reps=args.reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    if "k1" in only:
        n = 3000 if args.quick else 20000
        coords, _, targets = make_workload(n, 8833 if not args.quick else 1000, seed=0)
        cd = [ops.coords_to_device(c) for c in coords]
        N = 2 * n
        buf = torch.empty((N, ops.padded_ld(N)), dtype=torch.float64, device="cuda")[:, :N]
        xy = [ops.coords_to_device(np.random.default_rng(k).uniform(0, 1, (n, 2))) for k in (1, 2)]
        pe = list(HALF)
        pe[5:8] = [0.05, 0.05, 0.05]
        pg = list(GENERIC)
        pg[5:8] = [0.05, 0.05, 0.05]
        for name, c, p, metric in (("haversine_half", cd, HALF, METRIC_HAVERSINE), ("haversine_generic", cd, GENERIC, METRIC_HAVERSINE),
                                   ("euclid_half", xy, pe, METRIC_EUCLID), ("euclid_generic", xy, pg, METRIC_EUCLID)):
            ms = timed(lambda: ops.joint_cov(c, p, 2, metric, out=buf))
            res[f"k1_{name}"] = {"N": N, "ms": ms, "entries_per_s": N * N / ms * 1e3, "GBs_written": 8.0 * N * N / ms / 1e6,
                                 "unique_evals_per_s": (n * (n + 1) + n * n) / ms * 1e3}
        pd_ = ops.coords_to_device(targets)
        ms = timed(lambda: ops.cross_cov(cd, pd_, HALF, 2, 0, METRIC_HAVERSINE))
        res["k1_cross_haversine_half"] = {"N": N, "m": len(targets), "ms": ms, "GBs_written": 8.0 * N * len(targets) / ms / 1e6}
        del buf

    if "k2" in only:
        n = 2500 if args.quick else 10000
        lat, lon = np.arange(22.025, 58, 0.05), np.arange(-124.975, -65, 0.05)

        def draw(seed, k):
            idx = np.random.default_rng(seed).choice(len(lat) * len(lon), k, replace=False)
            return np.c_[lat[idx % len(lat)], lon[idx // len(lat)]]
        coords = [draw(2, n), draw(3, n)]
        values = [np.random.default_rng(5 + k).standard_normal(n) for k in range(2)]
        mf = fields.MultiField.from_arrays(coords, values, type="real")
        cfg = fields.VarioConfig(1500.0, 50, fast_dist=True)
        import time
        import warnings
        warnings.simplefilter("ignore")
        mf.empirical_variograms(cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev = mf.empirical_variograms(cfg)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        pairs_visited = n * (n - 1) // 2 * 2 + n * n
        res["k2_empirical_variograms_api"] = {"n_per_var": n, "bins": 50, "wall_ms": 1e3 * wall, "pairs_visited": pairs_visited,
                                              "pairs_per_s": pairs_visited / wall, "pairs_binned": int(ev.df["bin_count"].sum())}
        # kernel-only: pass 1 and pass 2 of the cross variogram (all n^2 pairs)
        Xa, Xb = ops.coords_to_device(coords[0]), ops.coords_to_device(coords[1])
        va, vb = ops.to_device(values[0]), ops.to_device(values[1])
        r = ops.vario_extrema(Xa, Xb, METRIC_HAVERSINE, False, 1500.0)
        centers, edges = fields._bins_from_extrema(r["min"], r["max"], 50)
        out = torch.empty(3, dtype=torch.float64, device="cuda")
        ws1 = torch.empty(int(_lib.lib.ck_vario_minmax_workspace_bytes(n, n)) // 8 + 1, dtype=torch.float64, device="cuda")
        ms1 = timed(lambda: _lib.check(_lib.lib.ck_vario_minmax(ops._ptr(Xa), n, ops._ptr(Xb), n, METRIC_HAVERSINE, 0, 1500.0, 0, -1,
                                                               ops._ptr(out), ops._ptr(ws1), ops._stream())))
        ms2 = timed(lambda: ops.vario_bin(Xa, va, 0.0, Xb, vb, 0.0, METRIC_HAVERSINE, False, False, 1500.0, edges))
        res["k2_cross_kernels"] = {"pairs": n * n, "minmax_ms": ms1, "bin_ms_incl_host_readback": ms2,
                                   "minmax_pairs_per_s": n * n / ms1 * 1e3, "bin_pairs_per_s": n * n / ms2 * 1e3,
                                   "effective_GBs_16B_per_pair": 16.0 * n * n / ms2 / 1e6}

    if "k3" in only:
        sizes = [2048, 4096] if args.quick else [4096, 8192, 16384, 32768, 40000]
        if args.k3_sizes:
            sizes = [int(v) for v in args.k3_sizes.split(",")]
        for N in sizes:
            xy = ops.coords_to_device(np.random.default_rng(4).uniform(0, 1, (N // 2, 2)))
            pe = [1, .8, 1.5, 1.5, 1.5, .05, .05, .05, .02, .02, -.2]
            buf = torch.empty((N, ops.padded_ld(N)), dtype=torch.float64, device="cuda")[:, :N]
            ws = ops.potrf_workspace(N, "cuda")
            t_a = timed(lambda: ops.joint_cov([xy, xy], pe, 2, METRIC_EUCLID, out=buf), reps=2)

            def fac():
                ops.joint_cov([xy, xy], pe, 2, METRIC_EUCLID, out=buf)
                return ops.potrf(buf, ws)
            t_f = timed(fac, reps=2) - t_a
            f = fac()
            m = 2048
            rb = torch.randn((m, ops.padded_ld(N)), dtype=torch.float64, device="cuda")[:, :N]
            t_s = timed(lambda: f.solve_lower(rb), reps=2)
            res[f"k3_N{N}"] = {"potrf_ms": t_f, "potrf_TFs": N ** 3 / 3 / t_f / 1e9, "info": f.info, "trsm_m": m, "trsm_ms": t_s,
                               "trsm_TFs": float(N) * N * m / t_s / 1e9}
            del buf, rb, f, ws
            torch.cuda.empty_cache()
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        res["cublas_dgemm_8192_TFs"] = 2 * 8192 ** 3 / timed(lambda: torch.matmul(a, b)) / 1e9
        del a, b

    if "k4" in only:
        import cokrig_b200  # noqa: F401
        params = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .01, .01, -.6]
        for name, nx, m, md in (("C1_40x40_500", 40, 500, 0.2), ("grid100_20000", 40 if args.quick else 100, 2000 if args.quick else 20000, 0.08)):
            gx = np.linspace(0, 1, nx)
            grid = np.array([(x, y) for y in gx for x in gx])
            z = [np.random.default_rng(k).standard_normal(len(grid)) for k in (1, 2)]
            pc = np.random.default_rng(7).uniform(0, 1, (m, 2))
            cd = [ops.coords_to_device(grid), ops.coords_to_device(grid)]
            zd = [ops.to_device(v) for v in z]
            pd_ = ops.coords_to_device(pc)
            out = {}

            def run():
                out["r"] = ops.local_predict(cd, zd, pd_, params, 2, 1, METRIC_EUCLID, md)
            ms = timed(run)
            k = out["r"][2]
            res[f"k4_{name}"] = {"targets": m, "ms_incl_readback": ms, "targets_per_s": m / ms * 1e3, "k_mean": float(k.mean()),
                                 "k_max": int(k.max()), "GFs": float(np.sum(k.astype(float) ** 3 / 3 + 2 * k.astype(float) ** 2)) / ms / 1e6}

    if "nll" in only:
        n = 3000 if args.quick else 20000
        coords, z, _ = make_workload(n, 10, seed=0)
        cd = [ops.coords_to_device(c) for c in coords]
        zd = ops.to_device(np.hstack(z))
        N = 2 * n
        buf = torch.empty((N, ops.padded_ld(N)), dtype=torch.float64, device="cuda")
        ws = ops.potrf_workspace(N, "cuda")
        for name, p in (("half", HALF), ("generic", GENERIC)):
            ms = timed(lambda: ops.gaussian_nll(cd, zd, p, 2, METRIC_HAVERSINE, sigma_buf=buf, ws=ws), reps=2)
            res[f"nll_{name}"] = {"N": N, "ms_per_eval": ms, "evals_per_s": 1e3 / ms, "TFs": (N ** 3 / 3 + 2.0 * N * N) / ms / 1e9}

    res["launches"] = int(_lib.lib.ck_launch_count())
    print(json.dumps(res, indent=1))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
