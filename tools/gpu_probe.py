"""Exploratory GPU probe: solve statistics and kernel timings (development tool)."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate, _lib
from simetuc_b200.settings import parse_power_dependence

def onp_tsim(cte):
    return float(np.round(15 * max(precalculate.get_lifetimes(cte)), 8))

def sweep(N_uc, n_pow, S=5.0, A=0.5, pulsed=True, nA=9):
    cte = configs.yb_tm_cte(N_uc=N_uc, S_conc=S, A_conc=A, activator_states=nA)
    t0 = time.time(); lat = lattice.generate(cte, seed=2001); tg = time.time() - t0
    t0 = time.time(); system = precalculate.build_system(cte, lat, n_sys=n_pow); tb = time.time() - t0
    powers = parse_power_dependence([1e6, 1e13, n_pow])
    for k, P in enumerate(powers):
        for exc in cte['excitations']['NIR_980']:
            exc.power_dens = P
        system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
    _, _, pop = system.ground_states()
    y0 = np.tile(pop.astype(float), (n_pow, 1))
    tf = onp_tsim(cte)
    print('N_uc %d ions %d N %d pairs %d stream %.1f MB lpr %d  gen %.2fs build %.2fs' % (
        N_uc, system.n_ions, system.N, system.n_pairs, system.stream_bytes / 1e6, system.lanes_per_row, tg, tb))
    if pulsed:
        t_pulse = np.linspace(0, 1e-8, 2)
        t_sol = np.logspace(np.log10(1e-8), np.log10(tf), 1000)
        t0 = time.time()
        yp, _, stp = system.integrate(t_pulse, y0, pulse=True, rtol=1e-8, atol=1e-15, want_y=True)
        ms_p, _ = system.last_timing()
        ys, avg, st = system.integrate(t_sol, yp[:, -1], pulse=False, rtol=1e-8, atol=1e-15, want_y=False, want_avg=True)
        ms_r, _ = system.last_timing()
        wall = time.time() - t0
        prof = system.last_profile(); tot_c = sum(v for k, v in prof.items() if k not in ('G', 'groups', 'arnoldi_stream')); print(' total Mcycles (sum over systems):', round(tot_c / 1e6, 1)); print(' profile:', {k: (round(v / tot_c, 3) if k not in ('G', 'groups') else v) for k, v in prof.items()})
        print(' pulse kernel %.1f ms relax kernel %.1f ms wall %.2f s -> %.1f solves/s' % (ms_p, ms_r, wall, n_pow / wall))
        for s in (stp[0], stp[-1], st[0], st[n_pow // 2], st[-1]):
            print('  ', {k: s[k] for k in ('n_steps', 'n_rejected', 'n_newton_fail', 'n_rhs', 'n_jv', 'status')})
        tot = sum(s['n_rhs'] + s['n_jv'] for s in st + stp)
        print(' field passes total %d -> %.2f us per pass per system-group' % (tot, (ms_p + ms_r) * 1e3 / max(tot, 1)))
    # rhs/jv kernel timing
    lib = system.lib
    import torch
    yd = torch.rand(n_pow, system.N, dtype=torch.float64, device='cuda')
    vd = torch.rand(n_pow, system.N, dtype=torch.float64, device='cuda')
    od = torch.empty_like(yd)
    ms = C.c_float()
    for nb in (1, n_pow):
        _lib.check(lib.stb_rhs_dev(system.handle, 0, nb, 1, yd.data_ptr(), od.data_ptr(), 20, C.byref(ms)))
        _lib.check(lib.stb_rhs_dev(system.handle, 0, nb, 1, yd.data_ptr(), od.data_ptr(), 50, C.byref(ms)))
        r = ms.value
        _lib.check(lib.stb_jv_dev(system.handle, 0, nb, 1, yd.data_ptr(), vd.data_ptr(), od.data_ptr(), 50, C.byref(ms)))
        j = ms.value
        b = system.stream_bytes + 16 * system.N
        print(' rhs x%d: %.4f ms  jv: %.4f ms   (stream %.2f MB per pass -> %.1f GB/s rhs per system)' % (
            nb, r, j, b / 1e6, nb * b / (r * 1e-3) / 1e9))
    system.close()

if __name__ == '__main__':
    N_uc = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    n_pow = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    S = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
    A = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
    pulsed = (sys.argv[5] != '0') if len(sys.argv) > 5 else True
    sweep(N_uc, n_pow, S, A, pulsed)
