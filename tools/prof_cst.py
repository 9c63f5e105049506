"""Sweep of the pair-pass cost model of the cluster integrator on the C2 batch (DEVELOPMENT TOOL)."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
import bench                                     # noqa: E402
cte, lat, system = bench.build_sweep_system(0, list(range(32)))
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
y0 = np.tile(pop.astype(float), (32, 1))
for base, slot, fix in ((2, 3, 6), (2, 3, 3), (2, 3, 0), (2, 3, 10), (6, 1, 0), (6, 1, 3), (8, 1, 0), (4, 2, 2), (10, 1, 0), (3, 2, 1)):
    os.environ['STB_CST_BASE'] = str(base)
    os.environ['STB_CST_SLOT'] = str(slot)
    os.environ['STB_PIECE_FIX'] = str(fix)
    best = (1e9, 1e9)
    for rep in range(2):
        _, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, y0, rtol=1e-8, atol=1e-15, want_y=False)
        ms = system.last_timing_dynamics()
        if sum(ms) < sum(best):
            best = ms
    print('fix %d, cst = %d + %d * slots: pulse %.2f relax %.2f ms -> %.0f solves/s' % (fix, base, slot, best[0], best[1], 32e3 / sum(best)), flush=True)
