"""Short integrator run for ncu (development tool): 4 systems of the C2 lattice, relaxation over a short span."""
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate
cte = configs.yb_tm_cte(N_uc=15)
lat = lattice.generate(cte, seed=2001)
n = 4
system = precalculate.build_system(cte, lat, n_sys=n)
_, _, pop = system.ground_states()
y0 = np.tile(pop.astype(float), (n, 1))
yp, _, st = system.integrate(np.linspace(0, 1e-8, 2), y0, pulse=True, rtol=1e-8, atol=1e-15)
t = np.logspace(-8, -5, 30)
ys, _, st = system.integrate(t, yp[:, -1], pulse=False, rtol=1e-8, atol=1e-15)
print(st[0], system.last_timing(), system.last_profile())
