"""Profiling driver: one factorisation of the config-4 matrix (m = 4M, n = 256) after one warm-up."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import enlsip_jl_b200 as E

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
n, nb = 256, 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(4)
W = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n)
truth = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
y = torch.tanh(W @ truth) + 0.01 * torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
x0 = truth.cpu().numpy() * 1.01
rho = (truth.cpu().numpy()[:4 * nb] ** 2).reshape(nb, 4).sum(axis=1)
mod = E.LargeCnlsModel("single_index", x0, {"W": W, "y": y, "rho": rho})
for i in range(reps):
    _, b, t = mod.factor(x0, want_R=False)
    print("build %.2f ms tsqr %.2f ms" % (b, t))
