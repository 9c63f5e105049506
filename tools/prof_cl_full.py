"""The bench's two integrator launches of the C2 batch (pulse, then the 1000-point relaxation) once,
for `ncu` (DEVELOPMENT TOOL)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench                                     # noqa: E402

cte, lat, system = bench.build_sweep_system(0, list(range(32)))
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
y0 = np.tile(pop.astype(float), (32, 1))
_, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, y0, rtol=1e-8, atol=1e-15, want_y=False)
print(system.last_timing_dynamics(), st[0]['n_steps'], st[31]['n_steps'])
