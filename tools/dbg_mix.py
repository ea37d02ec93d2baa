import sys, numpy as np
sys.path.insert(0,'.')
import enlsip_jl_b200 as E
from oracle import enlsip_oracle as O
from tests.test_large_user_family import mix_problem, MIX_SRC, MIX_M
pb,t,y,lo,up = mix_problem()
fam = E.LargeUserFamily(MIX_SRC, m=MIX_M, nb_ineqcons=1, data=("t", "y"), name="mix_user")
mod = E.LargeCnlsModel(fam, pb.x0, data={"t": t, "y": y}, x_low=lo, x_upp=up)
E.solve(mod, trace_cap=100)
print("engine", mod.exit_code, mod.iterations, mod.obj_value, mod.sol, mod.active)
for k in range(int(mod.iterations[0])+1):
    e = mod.trace[0][k]; print([float(v) for v in e[:14]], e[16:20])
r = O.solve(pb, wallclock=False)
print("oracle", r.exit_code, r.iterations, r.f, r.x, r.active)
for tr in r.trace: print(tr.t, tr.rankA, tr.rankJ2, tr.dimA, tr.dimJ2, tr.code, tr.alpha, tr.x_new)
R,_,_ = mod.factor(pb.x0)
J = pb.jac_res(pb.x0); rr = pb.res(pb.x0)
A = np.hstack([J, rr[:,None]])
Rn = np.linalg.qr(A, mode='r')
print("R dev", np.abs(np.abs(R)-np.abs(Rn)).max(), np.abs(Rn).max())
