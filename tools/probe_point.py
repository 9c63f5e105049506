"""One concentration point: which integrator runs and how long its launches take (DEVELOPMENT TOOL)."""
import os
import sys
import time
os.environ.setdefault('STB_VERBOSE', '1')
sys.path.insert(0, '.')
import numpy as np                                          # noqa: E402
from simetuc_b200 import configs, lattice, precalculate    # noqa: E402
import bench                                                # noqa: E402
S, A, nuc = float(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
cte = configs.yb_tm_cte(N_uc=nuc, S_conc=S, A_conc=A, power_dens=1e7, t_pulse=1e-8, rtol=1e-8, atol=1e-15, N_steps=1000)
t0 = time.perf_counter()
lat = lattice.generate(cte, seed=3000)
t1 = time.perf_counter()
system = precalculate.build_system(cte, lat, n_sys=1)
system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=0)
t2 = time.perf_counter()
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
yp, _, stp = system.integrate(t_pulse, pop.astype(float)[None, :], pulse=True, rtol=1e-8, atol=1e-15)
print('pulse', system.last_timing(), stp[0])
ys, avg, st = system.integrate(t_sol, yp[:, -1], pulse=False, rtol=1e-8, atol=1e-15, want_y=True, want_avg=True)
t3 = time.perf_counter()
print('relax', system.last_timing(), st[0])
print('ions %d N %d pairs %d | lattice %.3f s build %.3f s solve %.3f s' % (system.n_ions, system.N, system.n_pairs, t1 - t0, t2 - t1, t3 - t2))
