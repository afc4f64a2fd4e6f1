"""Phase timing of the diagonal-block kernel (ck_potf2_inv_kernel) from its clock64 stamps: python tools/potf2_phases.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sif-xco2-cokriging_b200"))
from cokrig_b200 import _lib, ops
n = 128
xy = np.random.default_rng(0).uniform(0, 1, (n, 2))
A = np.exp(-np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1)) / 0.2) + 0.01 * np.eye(n)
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
_lib.lib.ck_potf2_debug_buffer(dbg.data_ptr())
for it in range(3):
    a = torch.from_numpy(A.copy()).cuda()
    f = ops.potrf(a)
    torch.cuda.synchronize()
s = dbg.cpu().numpy()[:16].astype(np.int64)
names = ["load"] + [f"{w}{j}" for j in range(4) for w in ("elim", "panel", "trail")] + ["inverse", "store"]
d = np.diff(s)
for nm, v in zip(names, d):
    print(f"{nm:8s} {v:8d} cycles")
print("total", s[15] - s[0], "cycles")
_lib.lib.ck_potf2_debug_buffer(None)
