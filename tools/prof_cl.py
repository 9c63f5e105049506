"""C2 batch (32 powers) through the integrator with the per-phase cycle counters printed
(STB_VERBOSE=1).  DEVELOPMENT TOOL."""
import os
import sys
import numpy as np
os.environ.setdefault('STB_VERBOSE', '1')
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
import bench                                               # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cte = bench.make_cte()
powers = bench.powers_list()[:n] if n <= 32 else list(bench.powers_list()) * (n // 32)
n = len(powers)
lat = lattice.generate(cte, seed=2000)
system = precalculate.build_system(cte, lat, n_sys=n)
for k, P in enumerate(powers):
    for exc in cte['excitations']['NIR_980']:
        exc.power_dens = P
    system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
_, _, pop = system.ground_states()
y0 = np.tile(pop.astype(float), (n, 1))
t_pulse, t_sol = bench.t_grids(cte)
for rep in range(2):
    yp, _, stp = system.integrate(t_pulse, y0, pulse=True, rtol=1e-8, atol=1e-15)
    print('pulse ms', system.last_timing())
    ys, _, st = system.integrate(t_sol, yp[:, -1], pulse=False, rtol=1e-8, atol=1e-15, want_y=False, want_avg=True)
    print('relax ms', system.last_timing())
s0 = st[0]
print('system 0:', s0)
print('passes per system: min %d max %d' % (min(s['n_rhs'] + s['n_jv'] for s in st), max(s['n_rhs'] + s['n_jv'] for s in st)))
print('steps', [s['n_steps'] for s in st])
