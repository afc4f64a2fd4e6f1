"""GPU probe of the INT8 tensor-core update path (ck_oz_split / ck_oz_gemm): digit reconstruction, small and
ragged products against numpy, lower-triangular masking, and throughput.

    python tools/oz_probe.py [--perf] [--out gpurun_out/oz_probe.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sif-xco2-cokriging_b200"))
from cokrig_b200 import _lib  # noqa: E402

lib = _lib.lib
S = 7


def stream():
    return torch.cuda.current_stream().cuda_stream


def split(x: torch.Tensor, want_a=True, want_b=True):
    rows, k = x.shape
    fa = torch.zeros(lib.ck_oz_slices_bytes(rows, k, 0), dtype=torch.uint8, device="cuda") if want_a else None
    fb = torch.zeros(lib.ck_oz_slices_bytes(rows, k, 1), dtype=torch.uint8, device="cuda") if want_b else None
    sc = torch.zeros(lib.ck_oz_scales_len(rows), dtype=torch.float64, device="cuda")
    _lib.check(lib.ck_oz_split(x.data_ptr(), x.stride(0), rows, k, fa.data_ptr() if want_a else None,
                               fb.data_ptr() if want_b else None, sc.data_ptr(), stream()), "ck_oz_split")
    return fa, fb, sc


def reconstruct_a(fa, sc, rows, k):
    """A format [rb][kc][p][ku][rg16][r8][16] -> FP64 matrix."""
    kcn = k // 32
    rb = (rows + 127) // 128
    d = fa.cpu().numpy().view(np.int8).reshape(rb, kcn, S, 2, 16, 8, 16).astype(np.float64)
    w = 2.0 ** (-8.0 * np.arange(S))
    q = np.tensordot(d, w, axes=([2], [0]))  # rb kc ku rg r8 b
    q = q.transpose(0, 3, 4, 1, 2, 5).reshape(rb * 128, kcn * 32)
    return (q * sc.cpu().numpy()[:, None])[:rows]


def reconstruct_b(fb, sc, rows, k):
    """B format [hb][kc][ku][q][rg8][r8][16] -> FP64 matrix."""
    kcn = k // 32
    hb = (rows + 63) // 64
    d = fb.cpu().numpy().view(np.int8).reshape(hb, kcn, 2, S, 8, 8, 16).astype(np.float64)
    w = 2.0 ** (-8.0 * np.arange(S))
    q = np.tensordot(d, w, axes=([3], [0]))  # hb kc ku rg r8 b
    q = q.transpose(0, 3, 4, 1, 2, 5).reshape(hb * 64, kcn * 32)
    return (q * sc.cpu().numpy()[: hb * 64, None])[:rows]


def oz_gemm(a, b, c, lower=False):
    fa, _, sa = split(a, True, False)
    _, fb, sb = split(b, False, True)
    _lib.check(lib.ck_oz_gemm(fa.data_ptr(), sa.data_ptr(), a.shape[0], fb.data_ptr(), sb.data_ptr(), b.shape[0], a.shape[1],
                              c.data_ptr(), c.stride(0), 1 if lower else 0, 0, stream()), "ck_oz_gemm")
    torch.cuda.synchronize()


def check_case(m, n, k, lower, seed, wide_range=False):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((m, k))
    b = rng.standard_normal((n, k))
    if wide_range:
        a *= np.exp(rng.uniform(-30, 0, (m, k)))
        b *= np.exp(rng.uniform(-30, 0, (n, k)))
    c0 = rng.standard_normal((m, n))
    ref = c0 - a @ b.T
    if lower:
        keep = np.tril(np.ones((m, n), dtype=bool))
        ref = np.where(keep, ref, c0)
    ldc = n + (n % 2) + 2
    cbuf = torch.zeros((m, ldc), dtype=torch.float64, device="cuda")
    cbuf[:, :n] = torch.from_numpy(c0).cuda()
    cv = cbuf[:, :n]
    oz_gemm(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), cv, lower)
    got = cv.cpu().numpy()
    scale = (np.abs(a) @ np.abs(b).T).max()
    err = np.abs(got - ref)
    i, j = np.unravel_index(err.argmax(), err.shape)
    pad_ok = bool((cbuf[:, n:] == 0).all().item())
    return {"m": m, "n": n, "k": k, "lower": lower, "wide": wide_range, "max_err_rel": float(err.max() / scale),
            "argmax": [int(i), int(j)], "frac_bad": float((err > 1e-12 * scale).mean()), "pad_untouched": pad_ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--perf-only", action="store_true", help="skip the checks (profiling runs)")
    ap.add_argument("--sizes", default="16384x16384xL,32768x32768xL,8832x32768xR")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    out = {}

    if not args.perf_only:
        checks(out)
    if args.perf or args.perf_only:
        perf(out, args.sizes)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


def checks(out):
    # 1. digits reconstruct the panel to 2^-56 of the row maximum
    rng = np.random.default_rng(0)
    x = rng.standard_normal((200, 96)) * np.exp(rng.uniform(-20, 5, (200, 1)))
    x[7] = 0.0
    x[11, 5] = 0.98 * 2.0 ** 3
    xt = torch.from_numpy(x).cuda()
    fa, fb, sc = split(xt)
    torch.cuda.synchronize()
    ra, rb_ = reconstruct_a(fa, sc, 200, 96), reconstruct_b(fb, sc, 200, 96)
    rowmax = np.maximum(np.abs(x).max(axis=1, keepdims=True), 1e-300)
    out["split"] = {"a_err_rel_rowmax": float((np.abs(ra - x) / rowmax).max()), "b_err_rel_rowmax": float((np.abs(rb_ - x) / rowmax).max()),
                    "bound": 2.0 ** -55}
    print(json.dumps({"split": out["split"]}), flush=True)

    # 2. one tile, one chunk
    out["tile"] = check_case(128, 64, 32, False, 1)
    print(json.dumps({"tile": out["tile"]}), flush=True)

    # 3. deeper / ragged / masked cases
    out["cases"] = []
    for (m, n, k, lower, wide) in [(128, 64, 1024, False, False), (256, 192, 128, False, False), (1000, 900, 256, False, False),
                                   (1024, 1024, 1024, True, False), (1500, 1500, 512, True, True), (3000, 700, 1024, True, False),
                                   (2304, 5000, 1024, False, True)]:
        r = check_case(m, n, k, lower, 2 + m + n, wide)
        out["cases"].append(r)
        print(json.dumps(r), flush=True)



def perf(out, sizes):
    if True:
        out["perf"] = []
        for spec in sizes.split(","):
            ms, ns, ls = spec.split("x")
            m, n, lower = int(ms), int(ns), ls == "L"
            k = 1024
            a = torch.randn((m, k), dtype=torch.float64, device="cuda")
            b = a if lower else torch.randn((n, k), dtype=torch.float64, device="cuda")
            c = torch.randn((m, n), dtype=torch.float64, device="cuda")
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            fa, fb, sa = split(a)
            if not lower:
                _, fb, sb = split(b, False, True)
            else:
                sb = sa
            for _ in range(2):
                lib.ck_oz_gemm(fa.data_ptr(), sa.data_ptr(), m, fb.data_ptr(), sb.data_ptr(), n, k, c.data_ptr(), c.stride(0), int(lower), 0, stream())
            torch.cuda.synchronize()
            reps = 5
            e0.record()
            for _ in range(reps):
                lib.ck_oz_split(a.data_ptr(), a.stride(0), m, k, fa.data_ptr(), fb.data_ptr() if lower else None, sa.data_ptr(), stream())
            e1.record()
            for _ in range(reps):
                lib.ck_oz_gemm(fa.data_ptr(), sa.data_ptr(), m, fb.data_ptr(), sb.data_ptr(), n, k, c.data_ptr(), c.stride(0), int(lower), 0, stream())
            e2.record()
            torch.cuda.synchronize()
            t_split, t_gemm = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
            flops = 2.0 * m * n * k * (0.5 if lower else 1.0)
            # DMMA comparison
            cd = torch.randn((m, n), dtype=torch.float64, device="cuda")
            lib.ck_gemm_nt(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), cd.data_ptr(), cd.stride(0), m, n, k, 1, stream())
            torch.cuda.synchronize()
            e0.record()
            lib.ck_gemm_nt(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), cd.data_ptr(), cd.stride(0), m, n, k, 1, stream())
            e1.record()
            torch.cuda.synchronize()
            dbg = torch.zeros(8 * 148, dtype=torch.int64, device="cuda")
            lib.ck_oz_debug_buffer(dbg.data_ptr())
            lib.ck_oz_gemm(fa.data_ptr(), sa.data_ptr(), m, fb.data_ptr(), sb.data_ptr(), n, k, c.data_ptr(), c.stride(0), int(lower), 0, stream())
            torch.cuda.synchronize()
            lib.ck_oz_debug_buffer(None)
            d = dbg.cpu().numpy().reshape(148, 8).astype(float)
            tot = d[:, 0].mean()
            cyc = {"mma_total_cycles": tot, "mma_wait_operands_frac": d[:, 1].mean() / tot, "mma_wait_tmem_frac": d[:, 2].mean() / tot,
                   "epi_wait_frac": d[:, 4].mean() / tot, "epi_busy_frac": d[:, 5].mean() / tot}
            r = {"cycles": cyc, "m": m, "n": n, "k": k, "lower": lower, "split_ms": t_split, "oz_gemm_ms": t_gemm, "oz_TFs_fp64_equiv": flops / t_gemm / 1e9,
                 "dmma_full_rect_ms": e0.elapsed_time(e1), "dmma_TFs": 2.0 * m * n * k / e0.elapsed_time(e1) / 1e9}
            out["perf"].append(r)
            print(json.dumps(r), flush=True)
            del a, b, c, cd, fa, fb


if __name__ == "__main__":
    main()
