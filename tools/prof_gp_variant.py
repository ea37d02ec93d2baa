"""Development aid: time the batched C3 kernel of an alternative build of the library (path as argv[1])."""
import sys
sys.path.insert(0, '.')
import enlsip_jl_b200 as E
from enlsip_jl_b200.model import last_kernel_ms
if sys.argv[1] != "default":
    E.capi.LIB_PATH = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B)
m = E.CnlsModel('gauss_peaks', x0, data={'y': y, 'S': S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian='forward_diff')
for r in range(3):
    E.solve(m, want_active=False, want_counters=False)
print(sys.argv[1], B, "kernel ms %.2f" % last_kernel_ms(m), "solves/s %.4e" % (B / (last_kernel_ms(m) * 1e-3)), "converged", float((m.status_code == 1).mean()))
