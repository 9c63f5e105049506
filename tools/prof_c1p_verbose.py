"""Launch plan of the cluster integrator on the c1p / c3b lattices (STB_VERBOSE).  DEVELOPMENT TOOL."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
from tests.helpers import GOLDEN                          # noqa: E402
os.environ['STB_CLUSTER'] = '1'
os.environ['STB_VERBOSE'] = '1'
for name in ('c1p', 'c3b'):
    with np.load(os.path.join(GOLDEN, 'traj_%s.npz' % name)) as z:
        g = {k: z[k] for k in z.files}
    N_uc, S, A, seed, n_steps, power = g['params']
    cte = configs.yb_tm_cte(N_uc=int(N_uc), S_conc=float(S), A_conc=float(A), power_dens=float(power), t_pulse=1e-8,
                            rtol=1e-8, atol=1e-15, N_steps=int(n_steps))
    system = precalculate.build_system(cte, lattice.generate(cte, seed=int(seed)), n_sys=1)
    _, _, pop = system.ground_states()
    t_sol = g['t_sol']
    out = system.integrate_dynamics(np.linspace(0.0, t_sol[0], 2), t_sol, pop.astype(float), rtol=1e-8, atol=1e-15,
                                    nsteps=100000, want_y=False)
    print(name, 'ions', system.n_ions, 'ms', system.last_timing()[0], flush=True)
    system.close()
