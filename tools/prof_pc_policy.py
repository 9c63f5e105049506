"""Sweep of the preconditioner refresh policy of the cluster integrator on the C2 batch (DEVELOPMENT TOOL)."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
import bench                                     # noqa: E402

cte, lat, system = bench.build_sweep_system(0, list(range(32)))
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
y0 = np.tile(pop.astype(float), (32, 1))
for drift, age in ((0.3, 20), (0.5, 20), (1.0, 20), (2.0, 50), (0.15, 20), (0.3, 50), (1.0, 100)):
    os.environ['STB_PC_DRIFT'] = str(drift)
    os.environ['STB_PC_MAXAGE'] = str(age)
    best = (1e9, 1e9)
    for rep in range(2):
        _, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, y0, rtol=1e-8, atol=1e-15, want_y=False)
        ms = system.last_timing_dynamics()
        if sum(ms) < sum(best):
            best = ms
    print('drift %.2f maxage %3d: pulse %.2f relax %.2f ms -> %.0f solves/s | jv %d rhs %d nf %d rej %d ok %s' % (
        drift, age, best[0], best[1], 32e3 / sum(best), sum(s['n_jv'] for s in st), sum(s['n_rhs'] for s in st),
        sum(s['n_newton_fail'] for s in st), sum(s['n_rejected'] for s in st),
        all(s['status'] == 0 for s in st + st_p)), flush=True)
