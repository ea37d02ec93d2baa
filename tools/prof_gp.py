"""Profiling driver (development aid): one batched C3 solve with device-resident inputs."""
import sys, numpy as np
sys.path.insert(0, '.')
import enlsip_jl_b200 as E
from enlsip_jl_b200.model import last_kernel_ms
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
fam = sys.argv[2] if len(sys.argv) > 2 else 'gauss_peaks'
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if fam == 'hs65':
    x0 = E.synth.gen_hs65_batch(B); m = E.CnlsModel('hs65', x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
else:
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B)
    m = E.CnlsModel('gauss_peaks', x0, data={'y': y, 'S': S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian='forward_diff')
for r in range(reps):
    E.solve(m, want_active=False, want_counters=False)
    print(fam, B, "kernel ms", last_kernel_ms(m), "solves/s %.3e" % (B / (last_kernel_ms(m) * 1e-3)), m.kernel_info())
