"""Parity metric of the cluster integrator on the committed oracle trajectories for several GMRES tolerances
(max |delta| / (1e-4 |ref| + 1e-12) over the stored states; must stay <= 1).  DEVELOPMENT TOOL."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
from tests.helpers import GOLDEN, load_golden             # noqa: E402
from tests.test_gpu_parity import make_system             # noqa: E402

os.environ['STB_CLUSTER'] = '1'


def gold(name):
    with np.load(os.path.join(GOLDEN, 'traj_%s.npz' % name)) as z:
        return {k: z[k] for k in z.files}


def system_of(name, g):
    if name == 'fix240':
        return make_system(load_golden('240S_108A'), configs.sim_cte())
    N_uc, S, A, seed, n_steps, power = g['params']
    cte = configs.yb_tm_cte(N_uc=int(N_uc), S_conc=float(S), A_conc=float(A), power_dens=float(power), t_pulse=1e-8,
                            rtol=1e-8, atol=1e-15, N_steps=int(n_steps))
    return precalculate.build_system(cte, lattice.generate(cte, seed=int(seed)), n_sys=1)


for name in ('fix240', 'c1p', 'c3a', 'c3b'):
    g = gold(name)
    system = system_of(name, g)
    _, _, pop = system.ground_states()
    t_sol = g['t_sol']
    t_pulse = np.linspace(0.0, t_sol[0], 2)
    for lin in [float(v) for v in os.environ.get('LINS', '2e-3,1e-2,2e-2,5e-2,1e-1').split(',')]:
        os.environ['STB_LIN_RMS'] = str(lin)
        _, y, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=1e-8, atol=1e-15,
                                                        nsteps=100000, want_y=True)
        ms = system.last_timing()[0]
        ref = g['y_rows']
        m = float(np.max(np.abs(y[0][g['rows']] - ref) / (1e-4 * np.abs(ref) + 1e-12)))
        big = g['avg'] > 1e-9
        c = float(np.max(np.abs(avg[0] - g['avg'])[big] / g['avg'][big]))
        prof = system.last_profile()
        print('%-7s lin %.0e: %.1f ms (G=%d) steps %d jv %d rej %d | state metric %.3f, curves %.2e' % (
            name, lin, ms, prof['G'], st[0]['n_steps'], st[0]['n_jv'], st[0]['n_rejected'], m, c), flush=True)
    system.close()
