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
  N > 1 : ONE system of the same size factored and solved by all N GPUs together (strong scaling): the augmented
          array [Sigma ; C^T + z] in 1024-tiles, 2-D block-cyclic over a P x Q grid, panel broadcasts along
          process rows / all-gathers along process columns over NCCL (cokrig_b200.parallel.BlockCyclicCokriging;
          SURVEY 8e).  The independent-systems number (one system per GPU, no collective) is reported beside it
          under "replicas".
  --impl reference : the reference's CPU path (oracle port: scipy kv / sklearn haversine / LAPACK,
          all host threads -- also under torchrun, which exports OMP_NUM_THREADS=1) on a bounded sample of the
          same workload, scaled to the metric's unit.
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
_THREAD_LIMITS = None


def _all_host_threads() -> int:
    """Use every host core for BLAS / LAPACK even when the launcher (torchrun) exported OMP_NUM_THREADS=1."""
    global _THREAD_LIMITS
    cores = os.cpu_count() or 1
    try:
        os.sched_setaffinity(0, range(cores))
    except (AttributeError, OSError):
        pass
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cores)  # for libraries that are loaded after this point
    try:
        # the BLAS / LAPACK copies the CPU path uses must be LOADED before their thread pools can be resized:
        # numpy's and scipy's bundled OpenBLAS read OMP_NUM_THREADS=1 when they were imported
        import scipy.linalg  # noqa: F401
        import scipy.special  # noqa: F401
        import sklearn.metrics.pairwise  # noqa: F401
        import threadpoolctl
        _THREAD_LIMITS = threadpoolctl.threadpool_limits(limits=cores)  # kept alive: limits stay in force
    except Exception:  # noqa: BLE001
        pass
    return cores


def cpu_sample(n_s: int, m_s: int, N: int, m: int, n_asm: int = 4000) -> dict:
    """Reference CPU path (oracle port) on a bounded sample, scaled to the full workload by the O(N^2) entry count
    (assembly), O(N^3) (factorisations) and O(N^2 m) (solve) laws.  The assembly -- single-threaded scipy kv + sklearn
    haversine, the largest share of the reference's time -- is timed on its own, larger sample (one n_asm x n_asm
    marginal block and one cross block)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cokrig_oracle as orc
    coords, z, targets = make_workload(n_s, m_s, seed=0)
    P = orc.Params(PARAMS)
    t = orc.joint_predict_phases(P, I_PRED, coords, z, targets, "haversine")
    Ns = 2 * n_s
    ent_s, ent_f = Ns * Ns + Ns * m_s + m_s * m_s, N * N + N * m + m * m
    big, _, _ = make_workload(n_asm, 1, seed=1)
    t0 = time.perf_counter()
    orc.covariance(P, 0, orc.distance_matrix(big[0], big[0], fast_dist=True))
    orc.cross_covariance(P, 0, 1, orc.distance_matrix(big[0], big[1], fast_dist=True))
    asm_big_s = time.perf_counter() - t0
    per_entry = asm_big_s / (2.0 * n_asm * n_asm)
    scaled = {
        "assemble_s": per_entry * ent_f,
        "verify_s": t["verify_s"] * ((N + m) / (Ns + m_s)) ** 3,
        "factor_s": t["factor_s"] * (N / Ns) ** 3,
        "solve_s": t["solve_s"] * (2.0 * N * N * m + 2.0 * N * m * m) / (2.0 * Ns * Ns * m_s + 2.0 * Ns * m_s * m_s),
    }
    t_sample = sum(t[k] for k in scaled) + asm_big_s
    t_full = sum(scaled.values())
    return {"sample_s": t_sample, "sample_phases_s": {**{k: t[k] for k in scaled}, "assemble_large_s": asm_big_s},
            "scaled_step_s": t_full, "scaled_phases_s": scaled, "value": m / t_full,
            "value_without_verify": m / (t_full - scaled["verify_s"]),
            "assembly_entries_per_s": 1.0 / per_entry, "assembly_small_sample_entries_per_s": ent_s / t["assemble_s"],
            "sample": f"oracle port of src/joint_prediction.py:50-78 on n={n_s}/variable (N={Ns}), m={m_s} targets of the same "
                      f"lattice (verify / factor / solve), assembly on 2 x {n_asm}^2 entries; phase times scaled to N={N}, "
                      f"m={m} by entry count / N^3 / N^2 m; 'value' includes the reference's (m+N)^3/3 _verify_model "
                      "factorisation, 'value_without_verify' does not"}


def run_reference(args, N: int, m: int) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _all_host_threads()
    vals, last = [], None
    for it in range(args.warmup + args.steps):
        last = cpu_sample(args.cpu_n, args.cpu_m, N, m, args.cpu_asm_n)
        if it >= args.warmup:
            vals.append(last)
    step_s = statistics.mean(v["scaled_step_s"] for v in vals)
    value = m / step_s
    line = {"impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "predictions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3 joint cokriging, N={N} (n=20k/variable, 0.25deg CONUS lattice, haversine), m={m} targets, "
                                   "bivariate Matern nu=1.5", "inputs": "reference CPU path on a bounded sample, scaled"},
            "cpu_baseline": {"value": value, "unit": "predictions/s", "cores": cores, "kind": "port", "sample": last["sample"],
                             "sample_seconds": statistics.mean(v["sample_s"] for v in vals),
                             "value_without_verify": last["value_without_verify"], "scaled_phases_s": last["scaled_phases_s"],
                             "blas_threads": cores},
            "e2e": {"value": value, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def _ev():
    import torch
    return torch.cuda.Event(enable_timing=True)


def _timed(fn, reps=3, warm=1):
    """Best-of-reps duration in ms (CUDA events on the current stream, synchronised on both sides)."""
    import torch
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = _ev(), _ev()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def yardsticks(peaks: dict) -> dict:
    """Live yardsticks of this GPU, measured in this run with library GEMMs (calibration only): cuBLAS DGEMM 8192^3
    (FP64 tensor / vector peak -- MEASURED_PEAKS.json holds no FP64 figure) and a cuBLASLt int8 GEMM (the INT8
    tensor-core peak the update kernel is compared with)."""
    import torch
    out = {}
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    out["dgemm_TFs"] = 2 * 8192 ** 3 / _timed(lambda: torch.matmul(a, b), reps=4) / 1e9
    del a, b
    try:
        n = 16384
        ai = torch.randint(-100, 100, (n, n), dtype=torch.int8, device="cuda")
        bi = torch.randint(-100, 100, (n, n), dtype=torch.int8, device="cuda").t()  # column-major operand
        best = _timed(lambda: torch._int_mm(ai, bi), reps=5, warm=2)
        out["int8_TOPs"] = 2 * n ** 3 / best / 1e9
        # the same GEMM back to back for ~1.5 s: the rate the part sustains under its power cap
        torch.cuda.synchronize()
        e0, e1, reps = _ev(), _ev(), max(3, int(1500.0 / best))
        e0.record()
        for _ in range(reps):
            torch._int_mm(ai, bi)
        e1.record()
        torch.cuda.synchronize()
        out["int8_TOPs_sustained"] = 2 * n ** 3 * reps / e0.elapsed_time(e1) / 1e9
        out["int8_source"] = f"torch._int_mm (cuBLASLt int8 -> int32) {n}^3, best of 5 and {reps} back to back, measured live"
        del ai, bi
    except Exception as exc:  # noqa: BLE001
        out["int8_source"] = f"live int8 GEMM unavailable ({type(exc).__name__}); peak derived as 2 x bf16"
    bf16_sus, bf16_burst = peaks.get("bf16_tflops_sustained"), peaks.get("bf16_tflops")
    if not bf16_sus:
        bf16_sus, bf16_burst = 1400.0, 1590.0
        out["bf16_source"] = "fallback bf16 1.4 / 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
    else:
        out["bf16_source"] = "MEASURED_PEAKS.json"
    out["bf16_TFs_sustained"], out["bf16_TFs_burst"] = bf16_sus, bf16_burst
    return out


def update_kernel_roofline(achieved_int8: float, achieved_tf: float, fp64_flops: float, ys: dict, n_gpus: int, isolated) -> dict:
    """roofline object of the dominant kernel on the INT8 path.  Peak: the measured sustained bf16 GEMM rate of
    MEASURED_PEAKS.json x 2 (dense int8 = 2 x dense bf16 on sm_100a) -- the file has no int8 entry -- with the int8 GEMM
    measured live in this run reported beside it."""
    peak = 2 * ys["bf16_TFs_sustained"] * n_gpus
    r = {"bound": "tensor",
         "kernel": "ck_oz_gemm_kernel (tcgen05.mma kind::i8, TMEM accumulators): FP64-equivalent trailing / solve updates as 28 "
                   "int8 slice products per FP64 product",
         "achieved": achieved_int8, "peak": peak, "unit": "TFLOP/s", "frac": achieved_int8 / peak,
         "achieved_note": "int8 tensor ops (2 per MAC) per second over the factor + solve phases of the timed step, panel chains"
                          " (and, for N > 1, communication) included",
         "peak_source": f"{n_gpus} x 2 x bf16_tflops_sustained ({ys['bf16_source']}); int8 dense = 2 x bf16 dense on sm_100a",
         "fp64_equiv_TFs": achieved_tf, "dgemm_live_TFs": ys["dgemm_TFs"], "fp64_equiv_over_dgemm": achieved_tf / (n_gpus * ys["dgemm_TFs"]),
         # ncu --set full of the largest update of the factorisation (rows = 38976, lower, K = 1024;
         # profiles/r02x_ozgemm_l2hints3_ncu_full_summary.txt): dram read 19.15 GB + write 6.07 GB with the L2 eviction hints
         # (22.41 + 6.07 GB without them, profiles/r02x_ozgemm_l2hints0_ncu_full_summary.txt); algorithmic 12.7 GB
         "traffic": 25.22e9, "traffic_note": "per launch at rows=38976 (ncu r02x, CK_OZ_L2_HINTS=3: slices evict_last, C evict_first; "
                                             "28.48e9 without the hints), algorithmic 12.7e9 (C read + write 12.15e9, slices 0.56e9): "
                                             "operand slices re-read from DRAM (L2 hit rate 78 %); tensor pipe 80.3 % active",
         "flops_per_step": fp64_flops}
    if "int8_TOPs" in ys:
        r["int8_gemm_live_TOPs"] = {"burst": ys["int8_TOPs"], "sustained": ys["int8_TOPs_sustained"], "source": ys["int8_source"]}
        r["frac_of_live_int8_sustained"] = achieved_int8 / (n_gpus * ys["int8_TOPs_sustained"])
    if isolated:
        r["isolated_launch"] = isolated
    return r


def isolated_update_launch(N: int, ys: dict):
    """One launch of the largest update of the factorisation (rows = N - 1024, K = 1024) timed alone with CUDA events."""
    import torch
    from cokrig_b200 import _lib, ops
    rows, kk = N - 1024, 1024
    pa = torch.randn((rows, kk), dtype=torch.float64, device="cuda")
    cc = torch.zeros((rows, ops.padded_ld(rows)), dtype=torch.float64, device="cuda")
    fa = torch.empty(_lib.lib.ck_oz_slices_bytes(rows, kk, 0), dtype=torch.uint8, device="cuda")
    fb = torch.empty(_lib.lib.ck_oz_slices_bytes(rows, kk, 1), dtype=torch.uint8, device="cuda")
    sc = torch.empty(_lib.lib.ck_oz_scales_len(rows), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib.ck_oz_split(pa.data_ptr(), kk, rows, kk, fa.data_ptr(), fb.data_ptr(), sc.data_ptr(), st))
    ms = _timed(lambda: _lib.check(_lib.lib.ck_oz_gemm(fa.data_ptr(), sc.data_ptr(), rows, fb.data_ptr(), sc.data_ptr(), rows, kk,
                                                       cc.data_ptr(), cc.stride(0), 1, 0, st)), reps=3, warm=2)
    ops_launch = 28 * 2.0 * kk * rows * (rows + 1) / 2  # algorithmic: lower triangle only
    out = {"shape": f"lower update, rows={rows}, K={kk}", "ms": ms, "int8_TOPs": ops_launch / ms / 1e9,
           "frac_of_2x_bf16_burst": ops_launch / ms / 1e9 / (2 * ys["bf16_TFs_burst"]), "fp64_equiv_TFs": ops_launch / 28 / ms / 1e9}
    if "int8_TOPs" in ys:
        out["frac_of_live_int8_burst"] = ops_launch / ms / 1e9 / ys["int8_TOPs"]
    return out


def other_kernel_rooflines(ys: dict, peaks: dict, phase_ms: dict, wm: dict) -> dict:
    """Per-kernel roofline entries for the kernels that are not the dominant one, from CUDA-event times in this run:
    K1 (from the timed step), K2 and K4 on their own BASELINE configs (C2 cross-variogram, C1-like local batch)."""
    import torch
    from cokrig_b200 import METRIC_EUCLID, METRIC_HAVERSINE, ops
    hbm = peaks.get("hbm_gbs") or 6553.6
    out = {}
    asm_ms = phase_ms["assemble"] + phase_ms["cross"]
    gbs = wm["bytes_assembled"] / (asm_ms / 1e3) / 1e9
    out["K1 ck_block_kernel (covariance assembly, haversine + Matern 3/2)"] = {
        "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
        "entries_per_s": wm["entries"] / (asm_ms / 1e3),
        "note": "algorithmic bytes 8 (N^2 + N m) written; in practice FP64-pipe bound (ncu: profiles/r02*_k1_ncu*)"}
    # K2: C2 cross-variogram, 10 000 x 10 000 cells, 50 bins (pass 2 = the binning kernel)
    lat, lon = np.arange(22.025, 58, 0.05), np.arange(-124.975, -65, 0.05)

    def draw(seed, n):
        idx = np.random.default_rng(seed).choice(len(lat) * len(lon), n, replace=False)
        return np.c_[lat[idx % len(lat)], lon[idx // len(lat)]]
    ca, cb = draw(2, 10000), draw(3, 10000)
    Xa, Xb = ops.coords_to_device(ca), ops.coords_to_device(cb)
    va, vb = ops.to_device(np.random.default_rng(5).standard_normal(10000)), ops.to_device(np.random.default_rng(6).standard_normal(10000))
    ext = ops.vario_extrema(Xa, Xb, METRIC_HAVERSINE, False, 1500.0)
    centers = np.linspace(ext["min"], ext["max"], 50)
    w = centers[1] - centers[0]
    edges = np.arange(ext["min"] - 0.5 * w, ext["max"] + w, w)
    edges[0] = 0
    ms = _timed(lambda: ops.vario_bin(Xa, va, 0.0, Xb, vb, 0.0, METRIC_HAVERSINE, False, False, 1500.0, edges), reps=3)
    pairs = 1.0e8
    # FP64 yardstick for a vector-FP64-bound kernel: the FP64 pipe peak = live DGEMM rate / 2 flops per FMA lane-op
    out["K2 ck_vario_bin_kernel (pair binning, C2 cross 1e8 pairs, 50 bins)"] = {
        "bound": "fp64 pipe / shared-memory histogram", "achieved": pairs / (ms / 1e3) / 1e9, "unit": "Gpairs/s",
        "effective_GBs_at_16B_per_pair": 16 * pairs / (ms / 1e3) / 1e9, "ms_incl_readback": ms,
        "fp64_instr_per_pair": 60, "frac": 60 * pairs / (ms / 1e3) / (ys["dgemm_TFs"] * 1e12 / 2),
        "note": "no meaningful HBM traffic (O(n) reads): frac = FP64 instructions issued (~60 per haversine pair + bin search, "
                "ncu) / FP64 lane rate implied by the live DGEMM; includes the D2H of the 50 bins"}
    # K4: local-neighbourhood batch, 100 x 100 grid per variable, 5 000 targets, k ~ 370
    gx = np.linspace(0, 1, 100)
    grid = np.array([(x, y) for y in gx for x in gx])
    cd = [ops.coords_to_device(grid), ops.coords_to_device(grid)]
    zd = [ops.to_device(np.random.default_rng(k).standard_normal(len(grid))) for k in (1, 2)]
    pc = ops.coords_to_device(np.random.default_rng(7).uniform(0, 1, (5000, 2)))
    params = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .01, .01, -.6]
    sig = ops.joint_cov(cd, params, 2, METRIC_EUCLID)
    res = {}

    def run():
        res["k"] = ops.local_predict(cd, zd, pc, params, 2, 1, METRIC_EUCLID, 0.08, sigma=sig)[2]
    ms = _timed(run, reps=3)
    kf = res["k"].astype(float)
    tf = float(np.sum(kf ** 3 / 3 + 2 * kf ** 2)) / ms / 1e9
    out["K4 ck_local_predict_kernel (batched local systems, 5000 targets, k ~ 370)"] = {
        "bound": "tensor (FP64 DMMA)", "achieved": tf, "peak": ys["dgemm_TFs"], "unit": "TFLOP/s", "frac": tf / ys["dgemm_TFs"],
        "targets_per_s": 5000 / (ms / 1e3), "k_mean": float(kf.mean()), "ms_incl_count_pass_and_readback": ms,
        "note": "flops = k^3/3 + 2 k^2 per target; peak = live cuBLAS DGEMM"}
    del sig
    torch.cuda.empty_cache()
    return out


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
    strong = world > 1
    # N = 1: the rank's own system.  N > 1: ONE system (seed 0) shared by all ranks; the replicas leg uses per-rank systems.
    coords, z, targets = make_workload(n_per_var, m, seed=0)
    cd = [ops.coords_to_device(c) for c in coords]
    zd = ops.to_device(np.hstack(z))
    pd_ = ops.coords_to_device(targets)
    c0 = PARAMS[I_PRED] ** 2 + PARAMS[8 + I_PRED]
    ld = ops.padded_ld(N)
    phase_ms = {"assemble": 0.0, "potrf": 0.0, "cross": 0.0, "solve": 0.0}
    state = {}

    def step_single(record: bool):
        """One pass of the hot path on this GPU alone: K1 -> ck_potrf -> K1 -> ck_potrs_predict."""
        if "sigma" not in state:
            state["sigma"] = torch.empty((N, ld), dtype=torch.float64, device="cuda")[:, :N]
            state["ws"] = ops.potrf_workspace(N, "cuda")
        e = [_ev() for _ in range(5)]
        e[0].record()
        ops.joint_cov(cd, PARAMS, 2, METRIC_HAVERSINE, out=state["sigma"])
        e[1].record()
        f = ops.potrf(state["sigma"], state["ws"])
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

    solver = None
    if strong:
        from cokrig_b200 import parallel
        P_, Q_ = (int(v) for v in args.grid.lower().split("x")) if args.grid else (None, None)
        solver = parallel.BlockCyclicCokriging(parallel.ProcessGrid(P_, Q_), tile=args.tile, lookahead=True)

        def step(record: bool):
            pred, var, info = solver.solve_device(cd, zd, pd_, PARAMS, 2, I_PRED, METRIC_HAVERSINE)
            return (dict(solver._ev) if record else None), info, pred, var
    else:
        step = step_single

    for _ in range(args.warmup):
        _, f, pred, var = step(False)
    barrier()
    if strong:
        assert int(f.max().item()) == 0, "Sigma not positive definite"
    else:
        assert f.info == 0, "Sigma not positive definite"
    launches0 = _lib.lib.ck_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0, t1 = _ev(), _ev()
    recs = []
    barrier()
    t0.record()
    for _ in range(args.steps):
        recs.append(step(True)[0])
    t1.record()
    barrier()
    launches = _lib.lib.ck_launch_count() - launches0
    elapsed_ms = t0.elapsed_time(t1)
    if strong:
        phase_ms = {"assemble": 0.0, "factor_solve": 0.0, "reduce": 0.0}
        for e in recs:
            for name, a, b in (("assemble", "t0", "t1"), ("factor_solve", "t1", "t2"), ("reduce", "t2", "t3")):
                phase_ms[name] += e[a].elapsed_time(e[b]) / args.steps
    else:
        for e in recs:
            for k, name in enumerate(("assemble", "potrf", "cross", "solve")):
                phase_ms[name] += e[k].elapsed_time(e[k + 1]) / args.steps
    clocks = sampler.stop() if sampler else None
    pred_dev = pred.clone()
    if strong and os.environ.get("CK_MG_TRACE") and rank in (0, 1):
        rows = solver.trace_table()
        with open(os.path.join(ROOT, "gpurun_out", f"mg_trace_rank{rank}.json"), "w") as fh:
            json.dump(rows, fh)
    if args.profile:
        if rank == 0:
            print(f"profile run: {launches} launches in {args.steps} step(s), {elapsed_ms / args.steps:.1f} ms/step (not a bench value)",
                  file=sys.stderr)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public API with host buffers (H2D of coordinates / data / targets and D2H of the
    # predictions inside the timed region)
    import pandas as pd
    pframe = pd.DataFrame(targets, columns=["lat", "lon"])
    state.clear()
    torch.cuda.empty_cache()
    if strong:
        def e2e_call():
            p_, v_, info_ = solver.solve(coords, z, targets, PARAMS, 2, I_PRED, METRIC_HAVERSINE)
            assert info_ == 0
            return p_
        api = "cokrig_b200.parallel.BlockCyclicCokriging.solve (host numpy in, numpy out, every rank)"
    else:
        mf = fields.MultiField.from_arrays(coords, z, type="real")
        mod = model.MultivariateMatern(params=model.MaternParams().set_values(np.array(PARAMS)))
        predictor = joint_prediction.Predictor(mod, mf, fast_dist=True)

        def e2e_call():
            return predictor.predict_frame(I_PRED, pframe)["pred"].values
        api = "joint_prediction.Predictor.predict_frame (host numpy in, DataFrame out)"
    e2e_pred = e2e_call()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_pred = e2e_call()
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - w0)
    barrier()
    assert np.allclose(e2e_pred, pred_dev.cpu().numpy(), rtol=1e-9, atol=1e-12)
    h2d = 8 * (2 * N + N + 2 * m) * (world if strong else 1)  # strong: inputs are replicated on every rank
    d2h = (8 * 2 * m + 4) * (world if strong else 1)

    # ---- N > 1: the same sweep through the C handle API (ck_mg_create / ck_mg_joint_cov / ck_mg_potrf / ck_mg_potrs_predict:
    # schedule + NCCL row / column communicators behind the C ABI), device-resident inputs, max over ranks
    native = None
    if strong and not args.no_native:
        try:
            ns = parallel.NativeBlockCyclic(solver.g.P, solver.g.Q, tile=args.tile)
            for _ in range(2):
                pn, vn, info_n = ns.solve_device(cd, zd, pd_, PARAMS, 2, I_PRED, METRIC_HAVERSINE)
            barrier()
            n0_, n1_ = _ev(), _ev()
            n0_.record()
            for _ in range(args.steps):
                pn, vn, info_n = ns.solve_device(cd, zd, pd_, PARAMS, 2, I_PRED, METRIC_HAVERSINE)
            n1_.record()
            barrier()
            tn = torch.tensor([n0_.elapsed_time(n1_) / args.steps], dtype=torch.float64, device="cuda")
            dist.all_reduce(tn, op=dist.ReduceOp.MAX)
            native = {"value": m / (float(tn.item()) / 1e3), "unit": "predictions/s", "ms_per_step": float(tn.item()),
                      "info": int(info_n.item()),
                      "max_abs_pred_diff_vs_torch_distributed_sweep": float((pn - pred_dev).abs().max().item()),
                      "max_abs_var_diff_vs_torch_distributed_sweep": float((vn - var).abs().max().item()),
                      "local_GB": ns.local_bytes(N, m) / 1e9,
                      "api": "ck_mg_create / ck_mg_joint_cov / ck_mg_potrf / ck_mg_potrs_predict (include/cokrig.h, csrc/ck_mgctx.cu)"}
            assert native["info"] == 0 and native["max_abs_pred_diff_vs_torch_distributed_sweep"] < 1e-9 * float(pred_dev.abs().max().item())
            ns.close()
            del ns, pn, vn
        except Exception as exc:  # noqa: BLE001  (an extra leg: never lose the line over it)
            native = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

    # ---- N > 1: (a) the block-cyclic result against the single-GPU path (rank 0, untimed); (b) the independent-systems
    # number (one system per GPU, no collective) for reference
    replicas = None
    parity = None
    if strong and args.no_extras:
        rep_ms = float("nan")
    elif strong:
        del solver._keep
        torch.cuda.empty_cache()
        if rank == 0:
            _, f1, p1, v1 = step_single(False)
            # predictions of a zero-mean field cross zero: the solve error is normwise, so the gate is relative to the
            # largest prediction; the pointwise relative figure is reported beside it
            parity = {"max_abs_pred_diff_over_max_pred": float(((pred_dev - p1).abs().max() / p1.abs().max()).item()),
                      "max_rel_pred_vs_single_gpu": float(((pred_dev - p1).abs() / p1.abs().clamp_min(1e-300)).max().item()),
                      "max_abs_var_vs_single_gpu": float((var - v1).abs().max().item()), "info": f1.info}
            assert parity["max_abs_pred_diff_over_max_pred"] < 1e-9 and parity["max_abs_var_vs_single_gpu"] < 1e-9, parity
        barrier()
        coords_r, z_r, _ = make_workload(n_per_var, m, seed=rank)
        cd = [ops.coords_to_device(c) for c in coords_r]
        zd = ops.to_device(np.hstack(z_r))
        step_single(False)
        barrier()
        r0, r1 = _ev(), _ev()
        r0.record()
        for _ in range(2):
            step_single(False)
        r1.record()
        barrier()
        rep_ms = r0.elapsed_time(r1) / 2
        state.clear()
        torch.cuda.empty_cache()

    if world > 1:
        vals = [elapsed_ms, e2e_ms] + ([rep_ms] if strong else [])
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
        elapsed_ms, e2e_ms = vals[0], vals[1]
        if strong:
            replicas = None if args.no_extras else {"value": world * m / (vals[2] / 1e3), "unit": "predictions/s", "ms_per_step": vals[2], "scaling": "weak",
                        "what": "one independent C3 system per GPU (weekly-window semantics, SURVEY C4), no data-path collective"}
    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = m / (ms_per_step / 1e3) if strong else world * m / (ms_per_step / 1e3)
        wm = work_model(N, m)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        ys = yardsticks(peaks)
        # dominant kernel = the big trailing / solve updates, i.e. the potrf and solve phases (>= 96 % of the step):
        #   INT8 path (default): ck_oz_gemm_kernel, tcgen05 kind::i8, 28 int8 slice products per FP64 product
        #   CK_OZAKI=0         : ck_gemm_nt_kernel, FP64 DMMA
        gemm_s = (phase_ms["factor_solve"] if strong else phase_ms["potrf"] + phase_ms["solve"]) / 1e3
        fp64_flops = wm["potrf_flops"] + wm["solve_flops"]
        achieved_tf = fp64_flops / gemm_s / 1e12
        int8_active = bool(_lib.lib.ck_oz_active(N))
        if int8_active:
            roofline = update_kernel_roofline(28 * achieved_tf, achieved_tf, fp64_flops, ys, world,
                                              isolated_update_launch(N, ys) if not strong else None)
        else:
            roofline = {"bound": "tensor", "kernel": "ck_gemm_nt_kernel (FP64 DMMA: DSYRK trailing + TRSM updates)",
                        "achieved": achieved_tf, "peak": world * ys["dgemm_TFs"], "unit": "TFLOP/s",
                        "frac": achieved_tf / (world * ys["dgemm_TFs"]), "traffic": None,
                        "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (MEASURED_PEAKS.json holds no FP64 figure; "
                                       f"its hbm_gbs={peaks.get('hbm_gbs')})",
                        "flops_per_step": fp64_flops}
        line = {
            "metric": METRIC_NAME, "value": value, "unit": "predictions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3 joint cokriging, N={N} (n={n_per_var}/variable, 0.25deg CONUS lattice, haversine), m={m} "
                                   "targets, bivariate Matern nu=1.5" +
                                   (f"; ONE system on a {solver.g.P}x{solver.g.Q} block-cyclic grid of {world} GPUs, tile {args.tile}"
                                    if strong else ""),
                       "l2": "inputs larger than L2 (Sigma 8*N^2 bytes is rewritten every step)", "params": PARAMS},
            "phases_ms": phase_ms,
            "roofline": roofline,
            "update_path": "int8 tcgen05 (FP64-equivalent)" if int8_active else "fp64 dmma",
            "arithmetic": ("inputs, outputs, assembly, panel factorisation and all accumulation into Sigma / the right-hand sides in "
                           "f64; the rank-1024 trailing and solve updates as exact int8 x int8 -> int32 products of 7 balanced "
                           "base-256 digit slices of a 55-bit fixed-point rounding of the f64 operands (normwise: 2^-56 of the row "
                           "maximum; error below the f64 rounding of the same product; CK_OZAKI=0 runs them in f64 DMMA -- see "
                           "value_fp64_dmma)") if int8_active else "f64 throughout",
            "e2e": {"value": m / (e2e_ms / args.steps / 1e3) * (1 if strong else world), "unit": "predictions/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps, "api": api},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if strong:
            line["replicas"] = replicas
            line["parity_vs_single_gpu"] = parity
            line["native_handle_api"] = native
        else:
            line["assembly_GBs"] = wm["bytes_assembled"] / ((phase_ms["assemble"] + phase_ms["cross"]) / 1e3) / 1e9
            line["cholesky_TFs"] = wm["potrf_flops"] / (phase_ms["potrf"] / 1e3) / 1e12
            line["solve_TFs"] = wm["solve_flops"] / (phase_ms["solve"] / 1e3) / 1e12
        if world == 1:
            # the same step with every update in strict FP64 (DMMA), so that the two arithmetic paths are never conflated
            if int8_active and not args.no_dmma:
                _lib.lib.ck_oz_configure(0, -1)
                try:
                    step_single(False)
                    torch.cuda.synchronize()
                    d0, d1 = _ev(), _ev()
                    d0.record()
                    _, fd, pd64, _ = step_single(False)
                    d1.record()
                    torch.cuda.synchronize()
                    dm = d0.elapsed_time(d1)
                    line["value_fp64_dmma"] = {"value": m / (dm / 1e3), "unit": "predictions/s", "ms_per_step": dm,
                                               "fp64_TFs": fp64_flops / (dm / 1e3) / 1e12,
                                               "frac_of_dgemm_live": fp64_flops / (dm / 1e3) / 1e12 / ys["dgemm_TFs"],
                                               "max_rel_pred_int8_vs_dmma": float(((pd64 - pred_dev).abs() / pd64.abs().clamp_min(1e-300)).max().item()),
                                               "what": "CK_OZAKI=0: the whole step in strict FP64 (FP64 DMMA updates), 1 step"}
                finally:
                    _lib.lib.ck_oz_configure(1, -1)
                state.clear()
                torch.cuda.empty_cache()
            if not args.no_kernels:
                line["kernels"] = other_kernel_rooflines(ys, peaks, phase_ms, wm)
            if not args.no_cpu:
                cores = _all_host_threads()
                cb = cpu_sample(args.cpu_n, args.cpu_m, N, m, args.cpu_asm_n)
                line["cpu_baseline"] = {"value": cb["value"], "unit": "predictions/s", "cores": cores, "kind": "port",
                                        "sample": cb["sample"], "sample_seconds": cb["sample_s"],
                                        "value_without_verify": cb["value_without_verify"],
                                        "scaled_phases_s": cb["scaled_phases_s"],
                                        "assembly_entries_per_s": cb["assembly_entries_per_s"]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--cpu-asm-n", type=int, default=4000, help="points per variable of the CPU assembly sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-dmma", action="store_true", help="skip the strict-FP64 (CK_OZAKI=0) step")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline entries of K1 / K2 / K4")
    ap.add_argument("--no-native", action="store_true", help="N > 1: skip the leg through the C handle API (ck_mg_*)")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the single-GPU parity check and the replicas leg (tuning runs)")
    ap.add_argument("--grid", default="", help="process grid PxQ of the block-cyclic sweep (default: parallel.grid_shape)")
    ap.add_argument("--tile", type=int, default=1024, help="tile size of the block-cyclic sweep (N > 1)")
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
