"""Ownership regimes of the grid integrator (group blocks on / off, split pair pass, activator CTAs) over lattice
sizes and concentrations: short solves, curves with the group blocks against curves without.  DEVELOPMENT TOOL."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
os.environ['STB_CLUSTER'] = '0'
worst = 0.0
for n_uc, S, A in ((13, 25.0, 0.3), (14, 25.0, 2.0), (15, 40.0, 0.3), (16, 25.0, 0.3), (17, 10.0, 1.0), (18, 25.0, 0.1),
                   (19, 60.0, 0.3), (21, 25.0, 0.3), (22, 5.0, 0.5), (23, 25.0, 1.0), (26, 15.0, 0.3)):
    cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=S, A_conc=A, power_dens=1e8, t_pulse=1e-8, rtol=1e-8, atol=1e-15, N_steps=60)
    ref = None
    for tma, grp in (('0', '0'), ('0', '1'), ('1', '0'), ('1', '1')):
        os.environ['STB_TMA'] = tma
        os.environ['STB_GRID_GROUPS'] = grp
        system = precalculate.build_system(cte, lattice.generate(cte, seed=100 + n_uc), n_sys=1)
        _, _, pop = system.ground_states()
        t_sol = np.logspace(-8, -3.5, 60)
        _, _, avg, st_p, st = system.integrate_dynamics(np.linspace(0, 1e-8, 2), t_sol, pop.astype(float), rtol=1e-8, atol=1e-15,
                                                        nsteps=100000, want_y=False)
        ok = st_p[0]['status'] == 0 and st[0]['status'] == 0
        if ref is None:
            ref = avg[0]
        big = ref > 1e-9
        c = float(np.max(np.abs(avg[0] - ref)[big] / ref[big]))
        worst = max(worst, c)
        print('N_uc %2d S %4.1f A %3.1f (%5d ions) tma %s groups %s: %6.1f ms steps %3d jv %4d ok %s | vs first %.1e' % (
            n_uc, S, A, system.n_ions, tma, grp, system.last_timing()[0], st[0]['n_steps'], st[0]['n_jv'], ok, c), flush=True)
        assert ok
        system.close()
print('worst relative difference of the curves', worst)
assert worst < 1e-4
