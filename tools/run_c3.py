"""BASELINE config 3: concentration grid Yb {10,25,50} % x Tm {0.1,0.3,0.5,1,2} % x lattice
realisations, dynamics; lattices of 1000-1500 ions.  Times `simulate_concentration_dependence`
with 1 and with 8 lattices in flight per device.  MEASUREMENT TOOL."""
import json
import sys
import time
sys.path.insert(0, '.')
from simetuc_b200 import configs                   # noqa: E402
from simetuc_b200.simulations import Simulations   # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
concs, n_uc = [], []
for seed in range(n_seeds):
    for S, nuc in ((10.0, 20), (25.0, 15), (50.0, 12)):
        for A in (0.1, 0.3, 0.5, 1.0, 2.0):
            concs.append((S, A))
            n_uc.append(nuc)
out = {'points': len(concs)}
for streams in (1, 8, 16):
    cte = configs.yb_tm_cte(N_uc=15, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8,
                            rtol=1e-8, atol=1e-15, N_steps=1000)
    cte['lattice_seed'] = 3000
    sim = Simulations(cte, devices=[0], streams_per_device=streams)
    if streams == 1:
        sim.simulate_concentration_dependence(concs[:2], n_uc[:2], dynamics=True)     # warm-up
    t0 = time.perf_counter()
    sol = sim.simulate_concentration_dependence(concs, n_uc, dynamics=True)
    dt = time.perf_counter() - t0
    out['streams_%d' % streams] = {'seconds': dt, 'solves_per_s': len(concs) / dt}
    out['ions'] = [int(s.cte['ions']['total']) for s in sol.solution_list[:15]]
print(json.dumps(out))
