"""First-contact GPU check (development aid): engine vs oracle with traces on small batches."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import enlsip_jl_b200 as E
from enlsip_jl_b200.model import last_kernel_ms
from oracle import enlsip_oracle as O, problems as P
TR = 16

def cmp(fam, B, jac):
    if fam == 'hs65':
        x0 = E.synth.gen_hs65_batch(B); mk = lambda b: P.hs65(x0[b], fd=(jac == 'forward_diff'))
        m = E.CnlsModel('hs65', x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP, jacobian=jac)
    else:
        y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B); mk = lambda b: P.gauss_peaks(y[b], S[b], x0[b], fd=(jac == 'forward_diff'))
        m = E.CnlsModel('gauss_peaks', x0, data={'y': y, 'S': S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian=jac)
    print(fam, jac, m.kernel_info())
    E.solve(m, trace_cap=40)
    disc = 0; rels = []; relf = []; relpen = []
    for b in range(B):
        o = O.solve(mk(b), wallclock=False)
        act = [int(v) for v in m.active[b] if v > 0]
        ok = (o.exit_code == m.exit_code[b]) and (o.iterations == m.iterations[b]) and (o.active == act)
        if ok:
            for kk, tr in enumerate(o.trace[:40]):
                row = m.trace[b, kk]
                if (tr.t, tr.rankA, tr.rankJ2, tr.dimA, tr.dimJ2, tr.code, tr.index_del, tr.exit_code) != tuple(int(v) for v in (row[1], row[2], row[3], row[4], row[5], row[6], row[9], row[10])):
                    ok = False
        if not ok:
            disc += 1
            if disc <= 6: print("  DISCRETE MISMATCH", b, "oracle", o.exit_code, o.iterations, o.active, "engine", m.exit_code[b], m.iterations[b], act, o.threw)
            continue
        rels.append(np.linalg.norm(o.x - m.sol[b]) / np.linalg.norm(o.x)); relf.append(abs(o.f - m.obj_value[b]) / max(abs(o.f), 1e-300))
        if len(o.trace) >= 2:
            xo = o.trace[-2].x_new; xp = m.trace[b, len(o.trace) - 2, TR:]
            relpen.append(np.linalg.norm(xo - xp) / np.linalg.norm(xo))
    rels = np.array(rels); relf = np.array(relf); relpen = np.array(relpen)
    print("  B=%d discrete mismatches %d | x rel max %.2e med %.2e | f rel max %.2e | penultimate x rel max %.2e med %.2e | kernel %.3f ms"
          % (B, disc, rels.max(), np.median(rels), relf.max(), relpen.max(), np.median(relpen), last_kernel_ms(m)))

if __name__ == '__main__':
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cmp('hs65', B, 'analytic'); cmp('hs65', B, 'forward_diff'); cmp('gauss_peaks', B, 'analytic'); cmp('gauss_peaks', B, 'forward_diff')
    # throughput probe
    for fam, Bb in (('hs65', 200000), ('gauss_peaks', 200000)):
        if fam == 'hs65':
            x0 = E.synth.gen_hs65_batch(Bb); m = E.CnlsModel('hs65', x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
        else:
            y, S, x0, _ = E.synth.gen_gauss_peaks_batch(Bb); m = E.CnlsModel('gauss_peaks', x0, data={'y': y, 'S': S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian='forward_diff')
        for rep in range(3):
            t0 = time.time(); E.solve(m, want_active=False, want_counters=False); t1 = time.time()
            print("  %s B=%d wall %.3f s kernel %.3f ms -> %.3e solves/s (kernel)" % (fam, Bb, t1 - t0, last_kernel_ms(m), Bb / (last_kernel_ms(m) * 1e-3)))
        ec = np.asarray(m.exit_code); import collections
        print("  exit codes", dict(collections.Counter(ec.tolist()).most_common(8)), "mean iters %.2f" % np.mean(m.iterations))
