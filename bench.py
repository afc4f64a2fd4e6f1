"""bench.py -- headline benchmark of the cokriging hot path (BASELINE.json metric / config).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric: cokriging predictions/sec at n = 20k per variable (BASELINE.json configs[2], SURVEY 8d "C3"):
N = 40 000 stacked data on the 0.25 degree CONUS lattice (haversine km, bivariate Matern nu = 3/2),
m = 8 833 prediction targets (0.5 degree grid, 121 x 73).  One step = one full pass of the hot path:
assemble Sigma (K1) -> Cholesky (K3a) -> assemble C_dp (K1) -> triangular solve + prediction and
variance (K3b).  Inputs (12.8 GB Sigma, 2.8 GB C_dp) are far larger than L2, so no flush is needed.

  value : whole-job predictions/s with coordinates / data resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the reference-facing call joint_prediction.Predictor.predict_frame with
          HOST numpy buffers (H2D of coordinates + data, D2H of pred / pred_err inside the timed region)
  N > 1 : every rank solves its own independent system of the same size (weekly windows, SURVEY C4
          semantics): weak scaling, no data-path collective; NCCL only for the barrier / max-reduce.
  --impl reference : the reference's CPU path (oracle port: scipy kv / sklearn haversine / LAPACK,
          all host threads) on a bounded sample of the same workload, scaled to the metric's unit.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sif-xco2-cokriging_b200")
for _p in (os.path.join(PKG, "src"), PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

PARAMS = [1.0, 0.8, 1.5, 1.5, 1.5, 500.0, 500.0, 500.0, 0.02, 0.02, -0.2]  # SURVEY 8d half-integer set
I_PRED = 0
METRIC_NAME = "cokriging predictions/sec at n=20k"


# ------------------------------------------------------------------------------------------------ workload
def make_workload(n_per_var: int, m: int, seed: int):
    """points sampled from the 0.25 degree CONUS lattice (241 x 145 nodes, extents
    (-125, -65, 22, 58)); targets on the 0.5 degree grid (121 x 73 = 8 833, modelling_demo[10])."""
    lon = -125.0 + 0.25 * np.arange(241)
    lat = 22.0 + 0.25 * np.arange(145)
    nodes = np.array([(a, o) for a in lat for o in lon])
    coords = []
    for k in range(2):
        idx = np.random.default_rng(4 + k + 10 * seed).choice(len(nodes), n_per_var, replace=False)
        coords.append(np.ascontiguousarray(nodes[np.sort(idx)]))
    z = [np.random.default_rng(40 + k + 10 * seed).standard_normal(n_per_var) for k in range(2)]
    plon = -125.0 + 0.5 * np.arange(121)
    plat = 22.0 + 0.5 * np.arange(73)
    targets = np.array([(a, o) for a in plat for o in plon])
    if m < len(targets):
        targets = targets[np.sort(np.random.default_rng(7).choice(len(targets), m, replace=False))]
    return coords, z, np.ascontiguousarray(targets)


def work_model(N: int, m: int) -> dict:
    """Algorithmic work per step: entries assembled, Cholesky and solve flops."""
    return {"entries": N * N + N * m, "bytes_assembled": 8.0 * (N * N + N * m), "potrf_flops": N ** 3 / 3.0,
            "solve_flops": float(N) * N * (m + 1) + 2.0 * N * N}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, flag in zip(names, parts[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(n_s: int, m_s: int, N: int, m: int) -> dict:
    """Reference CPU path (oracle port) on a bounded sample, scaled to the full workload by the O(N^2)
    entry count (assembly), O(N^3) (factorisations) and O(N^2 m) (solve) laws."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cokrig_oracle as orc
    coords, z, targets = make_workload(n_s, m_s, seed=0)
    P = orc.Params(PARAMS)
    t = orc.joint_predict_phases(P, I_PRED, coords, z, targets, "haversine")
    Ns = 2 * n_s
    ent_s, ent_f = Ns * Ns + Ns * m_s + m_s * m_s, N * N + N * m + m * m
    scaled = {
        "assemble_s": t["assemble_s"] * ent_f / ent_s,
        "verify_s": t["verify_s"] * ((N + m) / (Ns + m_s)) ** 3,
        "factor_s": t["factor_s"] * (N / Ns) ** 3,
        "solve_s": t["solve_s"] * (2.0 * N * N * m + 2.0 * N * m * m) / (2.0 * Ns * Ns * m_s + 2.0 * Ns * m_s * m_s),
    }
    t_sample = sum(t[k] for k in scaled)
    t_full = sum(scaled.values())
    return {"sample_s": t_sample, "sample_phases_s": {k: t[k] for k in scaled}, "scaled_step_s": t_full,
            "scaled_phases_s": scaled, "value": m / t_full, "sample_value": m_s / t_sample,
            "sample": f"oracle port of src/joint_prediction.py:50-78 on n={n_s}/variable (N={Ns}), m={m_s} targets of the "
                      f"same lattice; phase times scaled to N={N}, m={m} by entry count / N^3 / N^2 m"}


def run_reference(args, N: int, m: int) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    vals, last = [], None
    for it in range(args.warmup + args.steps):
        last = cpu_sample(args.cpu_n, args.cpu_m, N, m)
        if it >= args.warmup:
            vals.append(last)
    step_s = statistics.mean(v["scaled_step_s"] for v in vals)
    value = m / step_s
    line = {"impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "predictions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3 joint cokriging, N={N} (n=20k/variable, 0.25deg CONUS lattice, haversine), m={m} targets, "
                                   "bivariate Matern nu=1.5", "inputs": "reference CPU path on a bounded sample, scaled"},
            "cpu_baseline": {"value": value, "unit": "predictions/s", "cores": cores, "kind": "port", "sample": last["sample"],
                             "sample_seconds": statistics.mean(v["sample_s"] for v in vals)},
            "e2e": {"value": value, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, n_per_var: int, m: int) -> None:
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: the NCCL version banner is written to fd 1 when the communicator is created,
        # so communicator creation (init + first collective) runs with fd 1 pointing at stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    from cokrig_b200 import METRIC_HAVERSINE, _lib, ops
    import fields, joint_prediction, model

    N = 2 * n_per_var
    coords, z, targets = make_workload(n_per_var, m, seed=rank)  # every rank owns its own system (window)
    cd = [ops.coords_to_device(c) for c in coords]
    zd = ops.to_device(np.hstack(z))
    pd_ = ops.coords_to_device(targets)
    c0 = PARAMS[I_PRED] ** 2 + PARAMS[8 + I_PRED]
    ld = ops.padded_ld(N)
    sigma = torch.empty((N, ld), dtype=torch.float64, device="cuda")[:, :N]
    ws = ops.potrf_workspace(N, "cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    phase_ms = {"assemble": 0.0, "potrf": 0.0, "cross": 0.0, "solve": 0.0}

    def step(record: bool):
        e = [ev() for _ in range(5)]
        e[0].record()
        ops.joint_cov(cd, PARAMS, 2, METRIC_HAVERSINE, out=sigma)
        e[1].record()
        f = ops.potrf(sigma, ws)
        e[2].record()
        cpd = ops.cross_cov(cd, pd_, PARAMS, 2, I_PRED, METRIC_HAVERSINE)
        e[3].record()
        pred, var = f.predict(cpd, zd, c0)
        e[4].record()
        return (e if record else None), f, pred, var

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        _, f, pred, var = step(False)
    barrier()
    assert f.info == 0, "Sigma not positive definite"
    launches0 = _lib.lib.ck_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0, t1 = ev(), ev()
    recs = []
    barrier()
    t0.record()
    for _ in range(args.steps):
        recs.append(step(True)[0])
    t1.record()
    barrier()
    launches = _lib.lib.ck_launch_count() - launches0
    elapsed_ms = t0.elapsed_time(t1)
    for e in recs:
        for k, name in enumerate(("assemble", "potrf", "cross", "solve")):
            phase_ms[name] += e[k].elapsed_time(e[k + 1]) / args.steps
    clocks = sampler.stop() if sampler else None
    pred_dev = pred.clone()
    if args.profile:
        if rank == 0:
            print(f"profile run: {launches} launches in {args.steps} step(s), {elapsed_ms / args.steps:.1f} ms/step (not a bench value)",
                  file=sys.stderr)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the drop-in API with host buffers
    mf = fields.MultiField.from_arrays(coords, z, type="real")
    mod = model.MultivariateMatern(params=model.MaternParams().set_values(np.array(PARAMS)))
    predictor = joint_prediction.Predictor(mod, mf, fast_dist=True)
    import pandas as pd
    pframe = pd.DataFrame(targets, columns=["lat", "lon"])
    del sigma, ws
    torch.cuda.empty_cache()
    for _ in range(min(args.warmup, 1) or 1):
        df = predictor.predict_frame(I_PRED, pframe)
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        df = predictor.predict_frame(I_PRED, pframe)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - w0)
    barrier()
    assert np.allclose(df["pred"].values, pred_dev.cpu().numpy(), rtol=1e-9, atol=1e-12)
    h2d = 8 * (2 * N + N + 2 * m)
    d2h = 8 * 2 * m + 4

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = t.tolist()
    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = world * m / (ms_per_step / 1e3)
        wm = work_model(N, m)
        # dominant kernel = the big trailing / solve updates, i.e. the potrf and solve phases (>= 96 % of the step):
        #   INT8 path (default): ck_oz_gemm_kernel, tcgen05 kind::i8, 28 int8 slice products per FP64 product
        #   CK_OZAKI=0         : ck_gemm_nt_kernel, FP64 DMMA
        gemm_s = (phase_ms["potrf"] + phase_ms["solve"]) / 1e3
        fp64_flops = wm["potrf_flops"] + wm["solve_flops"]
        achieved_tf = fp64_flops / gemm_s / 1e12
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        best = 1e9
        for _ in range(4):
            s0, s1 = ev(), ev()
            s0.record(); torch.matmul(a, b); s1.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        dgemm_tf = 2 * 8192 ** 3 / best / 1e9
        del a, b
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        int8_active = bool(_lib.lib.ck_oz_active(N))
        if int8_active:
            # isolated launch of the largest update of the factorisation (rows = N - 1024, K = 1024), CUDA events
            rows, kk = N - 1024, 1024
            pa = torch.randn((rows, kk), dtype=torch.float64, device="cuda")
            cc = torch.zeros((rows, ops.padded_ld(rows)), dtype=torch.float64, device="cuda")
            fa = torch.empty(_lib.lib.ck_oz_slices_bytes(rows, kk, 0), dtype=torch.uint8, device="cuda")
            fb = torch.empty(_lib.lib.ck_oz_slices_bytes(rows, kk, 1), dtype=torch.uint8, device="cuda")
            sc = torch.empty(_lib.lib.ck_oz_scales_len(rows), dtype=torch.float64, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib.ck_oz_split(pa.data_ptr(), kk, rows, kk, fa.data_ptr(), fb.data_ptr(), sc.data_ptr(), st))
            times = []
            for it in range(5):
                s0, s1 = ev(), ev()
                s0.record()
                _lib.check(_lib.lib.ck_oz_gemm(fa.data_ptr(), sc.data_ptr(), rows, fb.data_ptr(), sc.data_ptr(), rows, kk,
                                               cc.data_ptr(), cc.stride(0), 1, 0, st))
                s1.record(); torch.cuda.synchronize()
                if it >= 2:
                    times.append(s0.elapsed_time(s1))
            launch_ms = statistics.mean(times)
            int8_ops_launch = 28 * 2.0 * kk * rows * (rows + 1) / 2  # algorithmic: lower triangle only
            del pa, cc, fa, fb, sc
            bf16_sus, bf16_burst = peaks.get("bf16_tflops_sustained"), peaks.get("bf16_tflops")
            peak_src = "2 x MEASURED_PEAKS.json bf16_tflops_sustained (int8 dense = 2 x bf16 dense on sm_100a; no int8 figure in the file)"
            if not bf16_sus:
                bf16_sus, bf16_burst = 1400.0, 1590.0
                peak_src = "2 x fallback bf16 sustained 1.4 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
            achieved_int8 = 28 * fp64_flops / gemm_s / 1e12
            roofline = {
                "bound": "tensor",
                "kernel": "ck_oz_gemm_kernel (tcgen05.mma kind::i8, TMEM accumulators): FP64-equivalent trailing / solve updates as 28 "
                          "int8 slice products per FP64 product",
                "achieved": achieved_int8, "peak": 2 * bf16_sus, "unit": "TFLOP/s", "frac": achieved_int8 / (2 * bf16_sus),
                "achieved_note": "int8 tensor ops (2 per MAC) per second over the potrf + solve phases of the timed step, panel chains included",
                "peak_source": peak_src,
                "isolated_launch": {"shape": f"lower update, rows={rows}, K={kk}", "ms": launch_ms,
                                    "int8_TOPs": int8_ops_launch / launch_ms / 1e9,
                                    "frac_of_2x_bf16_burst": int8_ops_launch / launch_ms / 1e9 / (2 * bf16_burst),
                                    "fp64_equiv_TFs": int8_ops_launch / 28 / launch_ms / 1e9},
                "fp64_equiv_TFs": achieved_tf, "dgemm_live_TFs": dgemm_tf, "fp64_equiv_over_dgemm": achieved_tf / dgemm_tf,
                # ncu --set full of this launch shape (rows = 38976, lower, K = 1024; profiles/r01ai_ozgemm_ncu_full_summary.txt):
                # dram read 21.79 GB + write 6.08 GB; algorithmic 12.15 GB of C (read + write) + 0.56 GB of slices
                "traffic": 27.86e9, "traffic_note": "per launch at rows=38976 (ncu r01ai), algorithmic 12.7e9: the operand slices "
                                                    "are re-read ~14x from DRAM (L2 hit rate 76 %)",
                "flops_per_step": fp64_flops}
        else:
            roofline = {"bound": "tensor", "kernel": "ck_gemm_nt_kernel (FP64 DMMA: DSYRK trailing + TRSM updates)",
                        "achieved": achieved_tf, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": achieved_tf / dgemm_tf,
                        "traffic": None,
                        "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (MEASURED_PEAKS.json holds no FP64 figure; "
                                       f"its hbm_gbs={peaks.get('hbm_gbs')})",
                        "flops_per_step": fp64_flops}
        line = {
            "metric": METRIC_NAME, "value": value, "unit": "predictions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3 joint cokriging, N={N} (n={n_per_var}/variable, 0.25deg CONUS lattice, haversine), m={m} "
                                   "targets, bivariate Matern nu=1.5; one independent system per GPU",
                       "l2": "inputs larger than L2 (Sigma 8*N^2 bytes is rewritten every step)", "params": PARAMS},
            "phases_ms": phase_ms,
            "assembly_GBs": wm["bytes_assembled"] / ((phase_ms["assemble"] + phase_ms["cross"]) / 1e3) / 1e9,
            "cholesky_TFs": wm["potrf_flops"] / (phase_ms["potrf"] / 1e3) / 1e12,
            "solve_TFs": wm["solve_flops"] / (phase_ms["solve"] / 1e3) / 1e12,
            "roofline": roofline,
            "update_path": "int8 tcgen05 (FP64-equivalent)" if int8_active else "fp64 dmma",
            "arithmetic": ("inputs, outputs, assembly, panel factorisation and all accumulation into Sigma / the right-hand sides in "
                           "f64; the rank-1024 trailing and solve updates as exact int8 x int8 -> int32 products of 7 balanced "
                           "base-256 digit slices of a 55-bit fixed-point rounding of the f64 operands (error below the f64 "
                           "rounding of the same product; CK_OZAKI=0 runs them in f64 DMMA)") if int8_active else "f64 throughout",
            "e2e": {"value": world * m / (e2e_ms / args.steps / 1e3), "unit": "predictions/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "api": "joint_prediction.Predictor.predict_frame (host numpy in, DataFrame out)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cb = cpu_sample(args.cpu_n, args.cpu_m, N, m)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "predictions/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": cb["sample"], "sample_seconds": cb["sample_s"],
                                    "scaled_phases_s": cb["scaled_phases_s"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=20000, help="points per variable (default: the BASELINE config)")
    ap.add_argument("--m", type=int, default=8833, help="prediction targets")
    ap.add_argument("--cpu-n", type=int, default=3000, help="points per variable of the CPU sample")
    ap.add_argument("--cpu-m", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run under ncu (prints no bench line): 1 warm-up step, no e2e / cpu legs")
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = 1 if args.profile else max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, 2 * args.n, args.m)
    else:
        run_gpu(args, args.n, args.m)


if __name__ == "__main__":
    main()
