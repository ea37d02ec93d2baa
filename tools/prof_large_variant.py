"""Development aid: TSQR timing of an alternative build of the library (path as argv[1], 'default' = in-tree)."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import torch
import enlsip_jl_b200 as E
if sys.argv[1] != "default":
    E.capi.LIB_PATH = sys.argv[1]
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
n, nb = 256, 64
g = torch.Generator(device="cuda").manual_seed(4)
W = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n)
truth = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
y = torch.tanh(W @ truth) + 0.01 * torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
x0 = truth.cpu().numpy() * 1.01
rho = (truth.cpu().numpy()[:4 * nb] ** 2).reshape(nb, 4).sum(axis=1)
mod = E.LargeCnlsModel("single_index", x0, {"W": W, "y": y, "rho": rho})
for i in range(3):
    R, b, t = mod.factor(x0, want_R=(i == 2))
print(sys.argv[1], m, "build %.2f ms tsqr %.2f ms" % (b, t), "R checksum %.15g" % float(np.abs(R).sum()))
