"""Per-system step counts of the pulse and relaxation launches of the C2 batch (DEVELOPMENT TOOL)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench                                     # noqa: E402
cte, lat, system = bench.build_sweep_system(0, list(range(32)))
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
y0 = np.tile(pop.astype(float), (32, 1))
_, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, y0, rtol=1e-8, atol=1e-15, want_y=False)
print('kernel ms (pulse, relax):', system.last_timing_dynamics())
print('pulse steps:', [s['n_steps'] for s in st_p])
print('pulse passes:', [s['n_rhs'] + s['n_jv'] for s in st_p])
print('relax steps:', [s['n_steps'] for s in st])
print('relax passes:', [s['n_rhs'] + s['n_jv'] for s in st])
