"""One dynamics solve of the C1 lattice (N_uc from argv, default 30) through the grid integrator: the target
of the ncu capture of `k_integrate` (profiles/r02_ncu_integrate_grid.txt).  DEVELOPMENT TOOL."""
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
n_uc = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n_out = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8, rtol=1e-8, atol=1e-15,
                        N_steps=n_out)
system = precalculate.build_system(cte, lattice.generate(cte, seed=1000 + n_uc), n_sys=1)
_, _, pop = system.ground_states()
tf = float(np.round(15 * np.max(precalculate.get_lifetimes(cte)), 8))
t_sol = np.logspace(-8, np.log10(tf), n_out)
_, _, avg, st_p, st = system.integrate_dynamics(np.linspace(0, 1e-8, 2), t_sol, pop.astype(float), rtol=1e-8, atol=1e-15,
                                                nsteps=100000, want_y=False)
print(system.last_timing_dynamics(), st[0]['n_steps'], st[0]['n_rhs'], st[0]['n_jv'], system.n_pairs)
system.close()
