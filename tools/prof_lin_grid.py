"""GMRES tolerance of the grid integrator on C1-like lattices (DEVELOPMENT TOOL)."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
for n_uc in (20, 30):
    cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8, rtol=1e-8, atol=1e-15, N_steps=1000)
    lat = lattice.generate(cte, seed=1000 + n_uc)
    system = precalculate.build_system(cte, lat, n_sys=1)
    _, _, pop = system.ground_states()
    tf = float(np.round(15 * np.max(precalculate.get_lifetimes(cte)), 8))
    t_pulse = np.linspace(0, 1e-8, 2)
    t_sol = np.logspace(-8, np.log10(tf), 1000)
    ref = None
    for lin in (2e-3, 1e-2, 5e-2, 1e-1):
        os.environ['STB_LIN_RMS_GRID'] = str(lin)
        _, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=1e-8, atol=1e-15, nsteps=100000, want_y=False)
        ms = system.last_timing()[0]
        if ref is None:
            ref = avg[0]
        big = ref > 1e-9
        print('N_uc %d lin %.0e: %.1f ms steps %d rhs %d jv %d rej %d | curves vs lin 2e-3: %.2e' % (
            n_uc, lin, ms, st[0]['n_steps'], st[0]['n_rhs'], st[0]['n_jv'], st[0]['n_rejected'],
            float(np.max(np.abs(avg[0] - ref)[big] / ref[big]))), flush=True)
    system.close()
