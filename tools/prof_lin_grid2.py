"""GMRES tolerance of the grid integrator with the group preconditioner and the tightening rule: parity on
the committed trajectories (grid integrator forced) and time on C1 lattices.  DEVELOPMENT TOOL."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
from tests.helpers import GOLDEN, load_golden             # noqa: E402
from tests.test_gpu_parity import make_system             # noqa: E402
os.environ['STB_CLUSTER'] = '0'
LINS = [float(v) for v in os.environ.get('LINS', '2e-3,1e-2,2e-2,5e-2').split(',')]
for name in ('fix240', 'c1p', 'c3b'):
    with np.load(os.path.join(GOLDEN, 'traj_%s.npz' % name)) as z:
        g = {k: z[k] for k in z.files}
    if name == 'fix240':
        system = make_system(load_golden('240S_108A'), configs.sim_cte())
    else:
        N_uc, S, A, seed, n_steps, power = g['params']
        cte = configs.yb_tm_cte(N_uc=int(N_uc), S_conc=float(S), A_conc=float(A), power_dens=float(power), t_pulse=1e-8,
                                rtol=1e-8, atol=1e-15, N_steps=int(n_steps))
        system = precalculate.build_system(cte, lattice.generate(cte, seed=int(seed)), n_sys=1)
    _, _, pop = system.ground_states()
    t_sol = g['t_sol']
    for lin in LINS:
        os.environ['STB_LIN_RMS_GRID'] = str(lin)
        _, y, avg, st_p, st = system.integrate_dynamics(np.linspace(0.0, t_sol[0], 2), t_sol, pop.astype(float), rtol=1e-8,
                                                        atol=1e-15, nsteps=100000, want_y=True)
        ms = system.last_timing()[0]
        ref = g['y_rows']
        m = float(np.max(np.abs(y[0][g['rows']] - ref) / (1e-4 * np.abs(ref) + 1e-12)))
        big = g['avg'] > 1e-9
        c = float(np.max(np.abs(avg[0] - g['avg'])[big] / g['avg'][big]))
        print('%-7s lin %.0e: %.1f ms steps %d jv %d rej %d nf %d | state metric %.3f, curves %.2e' % (
            name, lin, ms, st[0]['n_steps'], st[0]['n_jv'], st[0]['n_rejected'], st[0]['n_newton_fail'], m, c), flush=True)
    system.close()
for n_uc in (20, 30):
    cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8, rtol=1e-8, atol=1e-15, N_steps=1000)
    system = precalculate.build_system(cte, lattice.generate(cte, seed=1000 + n_uc), n_sys=1)
    _, _, pop = system.ground_states()
    tf = float(np.round(15 * np.max(precalculate.get_lifetimes(cte)), 8))
    t_sol = np.logspace(-8, np.log10(tf), 1000)
    ref = None
    for lin in LINS:
        os.environ['STB_LIN_RMS_GRID'] = str(lin)
        _, _, avg, st_p, st = system.integrate_dynamics(np.linspace(0, 1e-8, 2), t_sol, pop.astype(float), rtol=1e-8, atol=1e-15,
                                                        nsteps=100000, want_y=False)
        ms = system.last_timing()[0]
        if ref is None:
            ref = avg[0]
        big = ref > 1e-9
        print('N_uc %d lin %.0e: %.1f ms steps %d rhs %d jv %d rej %d | curves vs first: %.2e' % (
            n_uc, lin, ms, st[0]['n_steps'], st[0]['n_rhs'], st[0]['n_jv'], st[0]['n_rejected'],
            float(np.max(np.abs(avg[0] - ref)[big] / ref[big]))), flush=True)
    system.close()
