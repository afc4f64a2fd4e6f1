"""Multi-GPU check (one process per GPU; run under torchrun on a box with >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/mg_check.py [--points 20000 --targets 8833 --tile 1024] [--out gpurun_out/mg_check.json]

1. K2: the row-block sharded variogram gives the SAME BITS (counts and FP64 sums) as the single-GPU call.
2. C5 path: BlockCyclicCokriging on the P x Q grid vs the single-GPU path (ck_potrf + ck_potrs_predict) and,
   at the small size, vs the CPU oracle (tests / tools may use oracle/ as the checker).
3. Timing of the block-cyclic solve at the requested size (device events, max over ranks).
Exit status 0 = all parity checks passed on every rank.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "sif-xco2-cokriging_b200", "src"), os.path.join(ROOT, "sif-xco2-cokriging_b200"),
           os.path.join(ROOT, "oracle"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

PARAMS = [1.0, 0.8, 1.5, 1.5, 1.5, 500.0, 500.0, 500.0, 0.02, 0.02, -0.2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=6000, help="points per variable of the timed block-cyclic solve")
    ap.add_argument("--targets", type=int, default=2000)
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--grid", default="", help="PxQ (default: as square as possible)")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--skip-single", action="store_true", help="do not run the single-GPU comparison at the timed size")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"],
                    help="c3: bench.py's CONUS workload; c5: 1 degree global lattice, 64 800 cells per variable (co-located), "
                         "targets = random cells, nugget 0.05 (BASELINE config 5; --points is ignored)")
    ap.add_argument("--sample-rows", type=int, default=12, help="rows of L gathered for the sampled L L^T check")
    ap.add_argument("--native", action="store_true", help="also time the C handle API (ck_mg_*) at the requested size")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import fields
    from bench import make_workload
    from cokrig_b200 import METRIC_HAVERSINE, ops, parallel

    report = {"world": world}
    ok = True

    # ---- 1. sharded variogram == single-GPU variogram, bit for bit
    lat, lon = np.arange(22.025, 58, 0.05), np.arange(-124.975, -65, 0.05)

    def draw(seed, n):
        idx = np.random.default_rng(seed).choice(len(lat) * len(lon), n, replace=False)
        return np.c_[lat[idx % len(lat)], lon[idx // len(lat)]]
    ca, cb = draw(2, 5000), draw(3, 4700)
    va, vb = np.random.default_rng(5).standard_normal(5000), np.random.default_rng(6).standard_normal(4700)
    shard = parallel.VarioShard()
    for same, (A, a, B, b) in ((True, (ca, va, ca, va)), (False, (ca, va, cb, vb))):
        one = fields._device_variogram(A, a, B, b, same, METRIC_HAVERSINE, False, 1500.0, 50)
        many = fields._device_variogram(A, a, B, b, same, METRIC_HAVERSINE, False, 1500.0, 50, shard=shard)
        same_bits = all(np.array_equal(x, y) for x, y in zip(one, many))
        report[f"vario_same_field_{same}_bit_identical"] = bool(same_bits)
        ok &= same_bits

    # ---- 2. block-cyclic cokriging vs single GPU vs oracle (small)
    P, Q = (int(v) for v in args.grid.split("x")) if args.grid else parallel.grid_shape(world)
    grid = parallel.ProcessGrid(P, Q)
    coords, z, targets = make_workload(1500, 700, seed=0)
    solver = parallel.BlockCyclicCokriging(grid, tile=256)
    pred, var, info = solver.solve(coords, z, targets, PARAMS, 2, 0, METRIC_HAVERSINE)
    cd = [ops.coords_to_device(c) for c in coords]
    f = ops.potrf(ops.joint_cov(cd, PARAMS, 2, METRIC_HAVERSINE))
    p1, v1 = f.predict(ops.cross_cov(cd, ops.coords_to_device(targets), PARAMS, 2, 0, METRIC_HAVERSINE),
                       ops.to_device(np.hstack(z)), PARAMS[0] ** 2 + PARAMS[8])
    p1, v1 = p1.cpu().numpy(), v1.cpu().numpy()
    e_pred = float(np.max(np.abs(pred - p1)) / np.max(np.abs(p1)))  # relative to the field scale (predictions cross zero)
    e_var = float(np.max(np.abs(var - v1)))
    e_ld = abs(solver.logdet() - float(f.logdet().item()))
    report.update({"small_pred_rel_vs_single_gpu": e_pred, "small_var_abs_vs_single_gpu": e_var, "small_info": info,
                   "small_logdet_abs": e_ld})
    ok &= e_pred < 1e-9 and e_var < 1e-9 and info == 0 and e_ld < 1e-7
    if rank == 0:
        import cokrig_oracle as orc
        rp, re, _ = orc.joint_predict(orc.Params(PARAMS), 0, coords, z, targets, "haversine")
        report["small_pred_rel_vs_oracle"] = float(np.max(np.abs(pred - rp)) / np.max(np.abs(rp)))
        report["small_var_abs_vs_oracle"] = float(np.max(np.abs(var - re ** 2)))
        ok &= report["small_pred_rel_vs_oracle"] < 1e-9 and report["small_var_abs_vs_oracle"] < 1e-9
    # the same system through the C handle API (csrc/ck_mgctx.cu: schedule + NCCL row / column communicators behind the ABI)
    native = parallel.NativeBlockCyclic(P, Q, tile=256)
    pn, vn, info_n = native.solve(coords, z, targets, PARAMS, 2, 0, METRIC_HAVERSINE)
    report.update({"native_small_pred_rel_vs_python_sweep": float(np.max(np.abs(pn - pred)) / np.max(np.abs(p1))),
                   "native_small_var_abs_vs_python_sweep": float(np.max(np.abs(vn - var))), "native_small_info": info_n,
                   "native_small_pred_rel_vs_single_gpu": float(np.max(np.abs(pn - p1)) / np.max(np.abs(p1))),
                   "native_small_logdet_abs": abs(native.logdet() - float(f.logdet().item()))})
    ok &= (report["native_small_pred_rel_vs_single_gpu"] < 1e-9 and float(np.max(np.abs(vn - v1))) < 1e-9 and info_n == 0
           and report["native_small_logdet_abs"] < 1e-7)
    native.close()
    del f, solver, native
    torch.cuda.empty_cache()

    # ---- 3. timed solve at the requested size
    params = list(PARAMS)
    if args.workload == "c5":
        glat, glon = -89.5 + np.arange(180.0), -179.5 + np.arange(360.0)
        cells = np.ascontiguousarray(np.array([(a, o) for a in glat for o in glon]))
        coords = [cells, cells.copy()]
        z = [np.random.default_rng(50 + k).standard_normal(len(cells)) for k in range(2)]
        targets = cells[np.sort(np.random.default_rng(7).choice(len(cells), args.targets, replace=False))]
        params[8] = params[9] = 0.05
        args.points = len(cells)
    else:
        coords, z, targets = make_workload(args.points, args.targets, seed=0)
    N = 2 * args.points
    solver = parallel.BlockCyclicCokriging(grid, tile=args.tile)
    report["local_GB"] = solver.local_bytes(N, len(targets)) / 1e9
    times = []
    for _ in range(args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pred, var, info = solver.solve(coords, z, targets, params, 2, 0, METRIC_HAVERSINE)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
        phases = dict(solver.timings)
    flops = N ** 3 / 3.0 + float(N) * N * (len(targets) + 1)
    if args.native:
        # the C handle API at the timed size: same inputs, device events around the three calls, max over ranks
        solver_py = solver
        native = parallel.NativeBlockCyclic(P, Q, tile=args.tile)
        ntimes = []
        for _ in range(args.steps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pn, vn, info_n = native.solve(coords, z, targets, params, 2, 0, METRIC_HAVERSINE)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ntimes.append(float(t.item()))
        report.update({"native_solve_ms": ntimes, "native_phases_ms_rank0": dict(native.timings), "native_info": info_n,
                       "native_pred_rel_vs_python_sweep": float(np.max(np.abs(pn - pred)) / np.max(np.abs(pred))),
                       "native_var_abs_vs_python_sweep": float(np.max(np.abs(vn - var))),
                       "native_local_GB": native.local_bytes(N, len(targets)) / 1e9})
        ok &= info_n == 0 and report["native_pred_rel_vs_python_sweep"] < 1e-9 and report["native_var_abs_vs_python_sweep"] < 1e-9
        native.close()
        del native
        torch.cuda.empty_cache()
    report.update({"workload": args.workload, "N": N, "m": len(targets), "tile": args.tile, "grid": f"{P}x{Q}", "solve_ms": times, "info": info,
                   "phases_ms_rank0": phases, "TFs_aggregate": flops / (min(times) / 1e3) / 1e12,
                   "predictions_per_s": len(targets) / (min(times) / 1e3)})
    ok &= info == 0
    # sampled reconstruction: rows of L gathered from their owners, (L L^T)[S, S] against Sigma[S, S] from the oracle
    if args.sample_rows > 0 and info == 0:
        S = np.sort(np.random.default_rng(11).choice(N, args.sample_rows, replace=False))
        S[-1] = N - 1
        Ls = solver.gather_factor_rows(S)
        if rank == 0:
            import cokrig_oracle as orc
            stacked = np.vstack(coords)
            P_ = orc.Params(params)
            ref = np.empty((len(S), len(S)))
            for a, ia in enumerate(S):
                for b, ib in enumerate(S):
                    pa_, pb_ = int(ia >= args.points), int(ib >= args.points)
                    d = orc.distance_matrix(stacked[ia:ia + 1], stacked[ib:ib + 1], fast_dist=True)
                    ref[a, b] = (orc.covariance(P_, pa_, d) if pa_ == pb_ else orc.cross_covariance(P_, min(pa_, pb_), max(pa_, pb_), d))[0, 0]
            rec = Ls @ Ls.T
            report["sampled_LLt_max_abs_err"] = float(np.max(np.abs(rec - ref)))
            report["sampled_rows"] = [int(v) for v in S]
            ok &= report["sampled_LLt_max_abs_err"] < 1e-9
        report["pred_finite"] = bool(np.isfinite(pred).all() and np.isfinite(var).all())
        report["var_min"], report["var_max"] = float(var.min()), float(var.max())
    if not args.skip_single and rank == 0:
        del solver
        torch.cuda.empty_cache()
        cd = [ops.coords_to_device(c) for c in coords]
        f = ops.potrf(ops.joint_cov(cd, params, 2, METRIC_HAVERSINE))
        p1, v1 = f.predict(ops.cross_cov(cd, ops.coords_to_device(targets), params, 2, 0, METRIC_HAVERSINE),
                           ops.to_device(np.hstack(z)), params[0] ** 2 + params[8])
        p1, v1 = p1.cpu().numpy(), v1.cpu().numpy()
        report["pred_rel_vs_single_gpu"] = float(np.max(np.abs(pred - p1)) / np.max(np.abs(p1)))
        report["var_abs_vs_single_gpu"] = float(np.max(np.abs(var - v1)))
        ok &= report["var_abs_vs_single_gpu"] < 1e-9

    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["ok"] = bool(flag.item() == 1.0)
    if rank == 0:
        print(json.dumps(report))
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(report, fh, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if report["ok"] else 1)


if __name__ == "__main__":
    main()
