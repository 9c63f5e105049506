"""Does want_y change the trajectory?  (DEVELOPMENT TOOL)"""
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
from simetuc_b200.simulations import Simulations          # noqa: E402

cte = configs.yb_tm_cte(N_uc=6, S_conc=10.0, A_conc=2.0, power_dens=1e6)
cte['lattice_seed'] = 21
sim = Simulations(cte)
d = sim.simulate_dynamics()
system = precalculate.build_system(cte, lattice.generate(cte, seed=21), n_sys=1)
_, _, pop = system.ground_states()
tf_p = sim._get_t_pulse()
t_pulse = np.linspace(0, tf_p, 2)
t_sol = np.logspace(np.log10(tf_p), np.log10(sim._get_t_simulation()), 1000)
res = []
for want_y in (True, False, True, False):
    for cum in (False, True):
        out = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=1e-8, atol=1e-15, want_y=want_y,
                                        want_avg=True, cumulative=cum)
        st = out[4][0]
        print(want_y, cum, st['n_steps'], st['n_jv'], st['n_rejected'], out[2][0][-1, 5:9], system.last_profile().get('G'))
        res.append(out[2][0])
print('avg equal TT vs FT (no cum):', np.array_equal(res[0], res[2]), np.max(np.abs(res[0] - res[2])))
print('cum equal:', np.array_equal(res[1], res[3]), np.max(np.abs(res[1] - res[3]) / np.maximum(np.abs(res[1]), 1e-300)))
y = d.y_sol
from scipy.integrate import cumulative_trapezoid as ct   # noqa: E402
import os                                                # noqa: E402
for lin in ('5e-2', '2e-3'):
    os.environ['STB_LIN_RMS'] = lin
    o1 = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=1e-8, atol=1e-15, want_y=True, want_avg=True)
    o2 = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=1e-8, atol=1e-15, want_y=False, want_avg=True,
                                   cumulative=True)
    y, avg, cum = o1[1][0], o1[2][0], o2[2][0]
    t0 = t_sol - t_sol[0]
    h_avg = ct(avg, t0, axis=0, initial=0)
    idx_S, idx_A, _ = system.ground_states()
    Y = ct(y, t0, axis=0, initial=0)
    cols = [Y[:, idx_A + j].mean(axis=1) for j in range(9)]
    print(lin, 'device cum  ', cum[-1, 2:])
    print(lin, 'host ct(avg)', h_avg[-1, 2:])
    print(lin, 'host avg(ct(y))', np.array([c[-1] for c in cols]))
    print(lin, 'min y', y.min(), 'min avg', avg.min(), 'state7 min/max over ions at row 500', y[500, idx_A + 7].min(), y[500, idx_A + 7].max())
