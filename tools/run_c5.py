"""BASELINE config 5: one optimiser iteration = 256 candidate ET-parameter sets on one lattice
(the C1 lattice at parity scale: N_uc=12, 25 % Yb, 0.3 % Tm, 9 Tm states), dynamics, evaluated as
one batch with `optimize.objective_batch`.  Also times config 1 (a single dynamics run on the same
lattice).  MEASUREMENT TOOL."""
import json
import os
import sys
import tempfile
import time
sys.path.insert(0, '.')
import numpy as np                                  # noqa: E402
from simetuc_b200 import configs, optimize          # noqa: E402
from simetuc_b200.simulations import Simulations    # noqa: E402

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cte = configs.yb_tm_cte(N_uc=12, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8, rtol=1e-8,
                        atol=1e-15, N_steps=1000)
cte['lattice_seed'] = 5000
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, 'expData', 'bNaYF4'))
t = np.linspace(0, 0.02, 400)
np.savetxt(os.path.join(tmp, 'expData', 'bNaYF4', 'decay_Tm_3H4_exc_NIR_980_25.0Yb_0.3Tm.txt'),
           np.column_stack((t, np.exp(-t / 1.5e-3) - np.exp(-t / 2e-4) + 0.01)))
os.chdir(tmp)
sim = Simulations(cte, devices=[0])
out = {}
t0 = time.perf_counter()
sol = sim.simulate_dynamics()
out['C1_first_call_s'] = time.perf_counter() - t0
t0 = time.perf_counter()
sol = sim.simulate_dynamics()
out['C1_dynamics_s'] = time.perf_counter() - t0
out['ions'] = int(sol.cte['ions']['total'])
out['states'] = int(sol.cte['states']['energy_states'])
base = [p.strength for p in cte['optimization']['processes']]
rng = np.random.default_rng(5000)
cands = [list(np.array(base) * 10.0 ** rng.uniform(-1, 1, len(base))) for _ in range(n_cand)]
errors, sols = optimize.objective_batch(sim, cands[:8])         # warm-up
t0 = time.perf_counter()
errors, sols = optimize.objective_batch(sim, cands)
dt = time.perf_counter() - t0
out['C5'] = {'candidates': n_cand, 'seconds': dt, 'solves_per_s': n_cand / dt,
             'all_converged': bool(all(s.solver_stats['relax']['status'] == 0 for s in sols)),
             'steps_min_max': [min(s.solver_stats['relax']['n_steps'] for s in sols),
                               max(s.solver_stats['relax']['n_steps'] for s in sols)]}
print(json.dumps(out))
