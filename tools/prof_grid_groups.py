"""Group block preconditioner of the grid integrator (STB_GRID_GROUPS=0/1): parity metric on the committed
oracle trajectories and time / pass counts on the C1 lattices.  DEVELOPMENT TOOL."""
import os
import sys
import time
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
from tests.helpers import GOLDEN, load_golden             # noqa: E402
from tests.test_gpu_parity import make_system             # noqa: E402

os.environ['STB_CLUSTER'] = '0'
big = [int(v) for v in os.environ.get('NUC', '20,30').split(',') if v]


def gold(name):
    with np.load(os.path.join(GOLDEN, 'traj_%s.npz' % name)) as z:
        return {k: z[k] for k in z.files}


def run(system, t_pulse, t_sol, pop, want_y):
    out = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=1e-8, atol=1e-15, nsteps=100000,
                                    want_y=want_y)
    return out, system.last_timing()[0]


for name in os.environ.get('CASES', 'c1p,c3b,fix240').split(','):
    if not name:
        continue
    g = gold(name)
    for grp in ('0', '1'):
        os.environ['STB_GRID_GROUPS'] = grp
        if name == 'fix240':
            system = make_system(load_golden('240S_108A'), configs.sim_cte())
        else:
            N_uc, S, A, seed, n_steps, power = g['params']
            cte = configs.yb_tm_cte(N_uc=int(N_uc), S_conc=float(S), A_conc=float(A), power_dens=float(power),
                                    t_pulse=1e-8, rtol=1e-8, atol=1e-15, N_steps=int(n_steps))
            system = precalculate.build_system(cte, lattice.generate(cte, seed=int(seed)), n_sys=1)
        _, _, pop = system.ground_states()
        t_sol = g['t_sol']
        (_, y, avg, st_p, st), ms = run(system, np.linspace(0.0, t_sol[0], 2), t_sol, pop, True)
        ref = g['y_rows']
        m = float(np.max(np.abs(y[0][g['rows']] - ref) / (1e-4 * np.abs(ref) + 1e-12)))
        bigv = g['avg'] > 1e-9
        c = float(np.max(np.abs(avg[0] - g['avg'])[bigv] / g['avg'][bigv]))
        print('%-7s groups %s: %.1f ms steps %d rhs %d jv %d rej %d nf %d status %d | state metric %.3f, curves %.2e' % (
            name, grp, ms, st[0]['n_steps'], st[0]['n_rhs'], st[0]['n_jv'], st[0]['n_rejected'], st[0]['n_newton_fail'],
            st[0]['status'], m, c), flush=True)
        system.close()

for n_uc in big:
    cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=float(os.environ.get('SCONC', '25')), A_conc=float(os.environ.get('ACONC', '0.3')), power_dens=float(os.environ.get('POWER', '1e7')), t_pulse=1e-8, rtol=1e-8, atol=1e-15,
                            N_steps=1000)
    ref_avg = None
    for grp in ('0', '1'):
        os.environ['STB_GRID_GROUPS'] = grp
        t0 = time.time()
        system = precalculate.build_system(cte, lattice.generate(cte, seed=1000 + n_uc), n_sys=1)
        t_build = time.time() - t0
        _, _, pop = system.ground_states()
        tf = float(np.round(15 * np.max(precalculate.get_lifetimes(cte)), 8))
        t_sol = np.logspace(-8, np.log10(tf), 1000)
        (_, _, avg, st_p, st), ms = run(system, np.linspace(0, 1e-8, 2), t_sol, pop, False)
        if ref_avg is None:
            ref_avg = avg[0]
        bigv = ref_avg > 1e-9
        c = float(np.max(np.abs(avg[0] - ref_avg)[bigv] / ref_avg[bigv]))
        print('C1 N_uc %d (%d ions) groups %s: %.1f ms (build %.2f s) steps %d rhs %d jv %d rej %d nf %d status %d | curves vs groups 0: %.2e' % (
            n_uc, system.n_ions, grp, ms, t_build, st[0]['n_steps'], st_p[0]['n_rhs'] + st[0]['n_rhs'],
            st_p[0]['n_jv'] + st[0]['n_jv'], st[0]['n_rejected'], st[0]['n_newton_fail'], st[0]['status'], c), flush=True)
        prof = system.last_profile()
        print('   cycles (M):', {k: round(v / 1e6, 1) for k, v in prof.items() if k not in ('G', 'groups')}, flush=True)
        system.close()
