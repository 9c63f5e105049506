"""One tiny solve through every kernel family (pair build, RHS / Jv, cluster integrator with and
without a cooperative process) - the target for compute-sanitizer runs.  DEVELOPMENT TOOL."""
import sys
sys.path.insert(0, '.')
import numpy as np                                          # noqa: E402
from simetuc_b200 import configs, lattice, precalculate    # noqa: E402
from simetuc_b200.engine import RateSystem                  # noqa: E402

coop = len(sys.argv) > 1 and sys.argv[1] == 'coop'
cte = configs.test_cte() if coop else configs.yb_tm_cte(N_uc=5, S_conc=10.0, A_conc=3.0, power_dens=1e9)
if coop:
    cte['lattice']['d_max_coop'] = np.inf
    cte['energy_transfer']['coop1'].strength = 5e11
    lat_cte = configs.yb_tm_cte(N_uc=5, S_conc=10.0, A_conc=3.0)
    arr = lattice.generate_arrays(lat_cte, seed=11)
    from simetuc_b200.lattice import Lattice
    lat = Lattice.from_arrays(arr['ion_type'], arr['dist_array'], 2, 7)
else:
    lat = lattice.generate(cte, seed=11)
system = RateSystem(lat, cte['energy_transfer'])
system.set_batch(3)
system.set_decay_blocks(*precalculate.decay_blocks(cte))
system.set_abs_blocks(*precalculate.absorption_blocks(cte))
_, _, pop = system.ground_states()
y0 = np.tile(pop.astype(float), (3, 1))
f = system.rhs(y0, pulse=True)
jv = system.jv(y0, f * 1e-9, pulse=True)
t = np.logspace(-8, -3, 25)
y, avg, st = system.integrate(t, y0, pulse=True, rtol=1e-6, atol=1e-15, want_avg=True)
print('ions', system.n_ions, 'N', system.N, 'coop triples', system.n_coop, [s['status'] for s in st], [s['n_steps'] for s in st],
      float(np.abs(y).max()))
system.close()
