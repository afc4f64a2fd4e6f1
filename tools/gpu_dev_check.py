"""Developer smoke/perf script run on the GPU box (not a pytest file). Writes gpurun_out/dev_check.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sif-xco2-cokriging_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import cokrig_oracle as orc
from cokrig_b200 import ops, METRIC_EUCLID, METRIC_HAVERSINE
from scipy.linalg import cholesky, solve_triangular

res = {}
def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den))

def timed(fn, reps=3):
    torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

rng = np.random.default_rng(0)
print(torch.cuda.get_device_name(0))
# ---- K1 parity
for name, params in [("half", [1, .8, 1.5, 1.5, 1.5, 500, 500, 500, .02, .02, -.2]),
                     ("generic", [1.0, 0.8, 0.75, 1.0, 1.25, 500, 500, 500, .02, .02, -.2]),
                     ("mixed", [1.1, 0.9, 0.5, 2.5, 3.5, 300, 400, 800, .01, .03, .3])]:
    lat = rng.uniform(22, 58, 700); lon = rng.uniform(-125, -65, 700)
    c0 = np.c_[lat[:400], lon[:400]]; c1 = np.c_[lat[300:], lon[300:]]   # partial co-location
    P = orc.Params(params)
    ref = orc.joint_cov(P, [c0, c1], "haversine")
    t = ops.joint_cov([ops.coords_to_device(c0), ops.coords_to_device(c1)], params, 2, METRIC_HAVERSINE)
    got = t.cpu().numpy()
    res[f"k1_joint_hav_{name}"] = rel(got, ref)
    pc = np.c_[rng.uniform(22, 58, 150), rng.uniform(-125, -65, 150)]; pc[:10] = c0[:10]
    refc = orc.pred_cross_cov(P, 1, [c0, c1], pc, "haversine")
    gotc = ops.cross_cov([ops.coords_to_device(c0), ops.coords_to_device(c1)], ops.coords_to_device(pc), params, 2, 1, METRIC_HAVERSINE)[:-1].cpu().numpy()
    res[f"k1_cross_hav_{name}"] = rel(gotc, refc.T)
    xy = rng.uniform(0, 1, (500, 2))
    pe = list(params); pe[5:8] = [.2, .25, .3]
    ref = orc.sim_joint_cov(orc.Params(pe), xy)
    d = ops.coords_to_device(xy)
    res[f"k1_joint_euc_{name}"] = rel(ops.joint_cov([d, d], pe, 2, METRIC_EUCLID).cpu().numpy(), ref)
h = np.concatenate([[0.0], 10 ** rng.uniform(-6, 4.3, 5000)])
for nu in [0.2, 0.5, 0.82, 1.5, 2.5, 3.2, 3.5]:
    ref = 1.3 ** 2 * orc.matern_correlation(nu, 500.0, h); ref[h == 0] += 0.02
    res[f"k1_eval_nu{nu}"] = rel(ops.matern_eval(h, 1.3 ** 2, nu, 500.0, 0.02), ref)
X = np.c_[rng.uniform(22, 58, 600), rng.uniform(-125, -65, 600)]
dref = orc.distance_matrix(X, X, fast_dist=True); dg = ops.distance_block(ops.coords_to_device(X), ops.coords_to_device(X), METRIC_HAVERSINE).cpu().numpy()
res["dist_hav_rel"] = rel(dg[dref > 0], dref[dref > 0]); res["dist_hav_biteq_frac"] = float((dg == dref).mean())
Y = rng.uniform(0, 1, (600, 2))
res["dist_euc_biteq"] = bool((ops.distance_block(ops.coords_to_device(Y), ops.coords_to_device(Y), METRIC_EUCLID).cpu().numpy() == orc.distance_matrix(Y, Y, units=None)).all())
print(json.dumps(res, indent=1)); sys.stdout.flush()

# ---- K3 parity: potrf vs scipy
for n in [100, 128, 129, 1000, 3200]:
    xy = rng.uniform(0, 1, (n, 2))
    A = np.exp(-orc.distance_matrix(xy, xy, units=None) / 0.2) + 0.01 * np.eye(n)
    Lref = cholesky(A, lower=True)
    buf = torch.empty((n, ops.padded_ld(n)), dtype=torch.float64, device="cuda")[:, :n]
    buf.copy_(torch.from_numpy(A))
    f = ops.potrf(buf)
    L = f.lower().cpu().numpy()
    res[f"potrf_n{n}_info"] = f.info
    res[f"potrf_n{n}_relL"] = float(np.abs(L - Lref).max() / np.abs(Lref).max())
    res[f"potrf_n{n}_resid"] = float(np.abs(L @ L.T - A).max())
    m = 37
    B = rng.standard_normal((m, n))
    Vref = solve_triangular(Lref, B.T, lower=True).T
    rb = torch.empty((m, ops.padded_ld(n)), dtype=torch.float64, device="cuda")[:, :n]; rb.copy_(torch.from_numpy(B))
    f.solve_lower(rb)
    res[f"trsm_n{n}_rel"] = float(np.abs(rb.cpu().numpy() - Vref).max() / np.abs(Vref).max())
# odd ld (unaligned path)
n = 301
xy = rng.uniform(0, 1, (n, 2)); A = np.exp(-orc.distance_matrix(xy, xy, units=None) / 0.2) + 0.01 * np.eye(n)
buf = torch.from_numpy(A.copy()).cuda()
f = ops.potrf(buf); L = f.lower().cpu().numpy()
res["potrf_oddld_resid"] = float(np.abs(L @ L.T - A).max()); res["potrf_oddld_info"] = f.info
# non-PD
A2 = A.copy(); A2[200, 200] = -1.0
f = ops.potrf(torch.from_numpy(A2).cuda()); res["potrf_nonpd_info"] = f.info
print(json.dumps({k: v for k, v in res.items() if k.startswith(("potrf", "trsm"))}, indent=1)); sys.stdout.flush()

# ---- joint prediction parity (C1-like)
params = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .01, .01, -.6]
P = orc.Params(params)
grid = orc.expand_grid(xcount=30, ycount=30)
_, _, fields = orc.sim_fields(P, grid, seed=1)
pc = np.random.default_rng(7).uniform(0, 1, (200, 2))
pr, pe, valid = orc.joint_predict(P, 1, [grid, grid], fields, pc, "euclidean")
cd = [ops.coords_to_device(grid), ops.coords_to_device(grid)]
z = ops.to_device(np.hstack(fields))
S = ops.joint_cov(cd, params, 2, METRIC_EUCLID)
f = ops.potrf(S)
cpd = ops.cross_cov(cd, ops.coords_to_device(pc), params, 2, 1, METRIC_EUCLID)
pred, var = f.predict(cpd, z, P.sigma[1, 1] ** 2 + P.nugget[1, 1])
res["joint_pred_rel"] = rel(pred.cpu().numpy(), pr)
res["joint_err_rel"] = rel(np.nan_to_num(np.sqrt(var.cpu().numpy())), pe)
res["joint_info"] = f.info
nll_ref = orc.gaussian_nll(P, [grid, grid], fields, "euclidean")
out, info = ops.gaussian_nll(cd, z, params, 2, METRIC_EUCLID)
res["nll_rel"] = rel(out.cpu().numpy()[0], nll_ref)
# ---- point prediction parity
pr2, sd2, k2, v2 = orc.point_predict(P, 1, [grid, grid], fields, pc[:60], 0.2, "euclidean")
g_pred, g_sd, g_k, g_info = ops.local_predict(cd, [ops.to_device(fields[0]), ops.to_device(fields[1])], ops.coords_to_device(pc[:60]), params, 2, 1, METRIC_EUCLID, 0.2)
res["local_k_equal"] = bool((g_k == k2).all()); res["local_pred_rel"] = rel(g_pred, pr2); res["local_sd_rel"] = rel(g_sd, sd2)
res["local_kmax"] = int(g_k.max())
# ---- variogram parity
lat = np.arange(22.025, 58, 0.05); lon = np.arange(-124.975, -65, 0.05)
def draw(seed, n):
    r = np.random.default_rng(seed); idx = r.choice(len(lat) * len(lon), n, replace=False)
    return np.c_[lat[idx % len(lat)], lon[idx // len(lat)]]
ca, cb = draw(2, 1500), draw(3, 1400)
va, vb = np.random.default_rng(5).standard_normal(1500), np.random.default_rng(6).standard_normal(1400)
for (i, j) in [(0, 0), (0, 1)]:
    ref = orc.get_variogram([ca, cb], [va, vb], i, j, 1500.0, 50, "haversine")
    A_, B_ = (ca, ca) if i == j else (ca, cb); a_, b_ = (va, va) if i == j else (va, vb)
    Xa, Xb = ops.coords_to_device(A_), ops.coords_to_device(B_)
    mn, mx, cnt = ops.vario_minmax(Xa, Xb, METRIC_HAVERSINE, i == j, 1500.0)
    centers = np.linspace(mn, mx, 50); w = centers[1] - centers[0]
    edges = np.arange(mn - 0.5 * w, mx + w, w); edges[0] = 0
    counts, sums = ops.vario_bin(Xa, ops.to_device(a_), a_.mean(), Xb, ops.to_device(b_), b_.mean(), METRIC_HAVERSINE, i == j, False, 1500.0, edges)
    res[f"vario{i}{j}_counts_equal"] = bool((counts == ref["bin_count"].values).all())
    res[f"vario{i}{j}_centers_equal"] = bool((centers == ref["bin_center"].values).all())
    res[f"vario{i}{j}_mean_rel"] = rel(sums / counts, ref["bin_mean"].values)
    res[f"vario{i}{j}_npairs"] = cnt
print(json.dumps(res, indent=1)); sys.stdout.flush()

# ---- performance
perf = {}
for n in [4096, 8192, 20000, 40000]:
    ld = ops.padded_ld(n)
    xy = torch.from_numpy(np.random.default_rng(4).uniform(0, 1, (n // 2, 2))).cuda()
    pe = [1, .8, 1.5, 1.5, 1.5, .05, .05, .05, .02, .02, -.2]
    buf = torch.empty((n, ld), dtype=torch.float64, device="cuda")[:, :n]
    ws = ops.potrf_workspace(n, "cuda")
    t_as = timed(lambda: ops.joint_cov([xy, xy], pe, 2, METRIC_EUCLID, out=buf), reps=2)
    def fac():
        ops.joint_cov([xy, xy], pe, 2, METRIC_EUCLID, out=buf)
        return ops.potrf(buf, ws)
    t_both = timed(fac, reps=2)
    f = fac(); inf = f.info
    t_f = t_both - t_as
    perf[f"n{n}"] = dict(assemble_ms=t_as, assemble_GBs=8 * n * n / t_as / 1e6, potrf_ms=t_f, potrf_TFs=n ** 3 / 3 / t_f / 1e9, info=inf)
    m = 2048
    rb = torch.randn((m, ld), dtype=torch.float64, device="cuda")[:, :n]
    t_s = timed(lambda: f.solve_lower(rb), reps=1)
    perf[f"n{n}"].update(trsm_m=m, trsm_ms=t_s, trsm_TFs=n * n * m / t_s / 1e9)
    print(n, perf[f"n{n}"]); sys.stdout.flush()
    del buf, rb, f, ws
    torch.cuda.empty_cache()
# dgemm yardstick
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
t = timed(lambda: torch.matmul(a, b), reps=3); perf["cublas_dgemm_8192_TFs"] = 2 * 8192 ** 3 / t / 1e9
t = timed(lambda: torch.linalg.cholesky(a @ a.T + 8192 * torch.eye(8192, dtype=torch.float64, device="cuda")), reps=2)
print("cusolver chol 8192 incl gemm ms", t)
res["perf"] = perf
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_check.json"), "w"), indent=1)
print(json.dumps(perf, indent=1))
