"""Pulse launch of the C2 batch for ncu (DEVELOPMENT TOOL)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import lattice, precalculate   # noqa: E402
import bench                                      # noqa: E402
n = 32
cte = bench.make_cte()
powers = bench.powers_list()
lat = lattice.generate(cte, seed=2000)
system = precalculate.build_system(cte, lat, n_sys=n)
for k, P in enumerate(powers):
    for exc in cte['excitations']['NIR_980']:
        exc.power_dens = P
    system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
_, _, pop = system.ground_states()
y0 = np.tile(pop.astype(float), (n, 1))
t_pulse, t_sol = bench.t_grids(cte)
yp, _, stp = system.integrate(t_pulse, y0, pulse=True, rtol=1e-8, atol=1e-15)
print(system.last_timing(), stp[0])
t = np.logspace(-8, -4.5, 40)
ys, _, st = system.integrate(t, yp[:, -1], pulse=False, rtol=1e-8, atol=1e-15, want_y=False, want_avg=True)
print(system.last_timing(), st[0], st[31])
