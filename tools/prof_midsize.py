"""Mid-size lattices (pair data between the shared memory of a cluster and the HBM-bound regime):
cluster integrator vs grid integrator with several group sizes.  DEVELOPMENT TOOL."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402

cases = [(12, 25.0, 0.3), (15, 25.0, 0.3), (20, 25.0, 0.3), (20, 10.0, 0.1), (12, 50.0, 2.0)]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for n_uc, S, A in cases:
    cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=S, A_conc=A, power_dens=1e7, t_pulse=1e-8, rtol=1e-8,
                            atol=1e-15, N_steps=1000)
    lat = lattice.generate(cte, seed=1000 + n_uc)
    system = precalculate.build_system(cte, lat, n_sys=1)
    _, _, pop = system.ground_states()
    tf = float(np.round(15 * np.max(precalculate.get_lifetimes(cte)), 8))
    t_pulse = np.linspace(0, 1e-8, 2)
    t_sol = np.logspace(-8, np.log10(tf), 1000)
    print('N_uc %d %g%%/%g%%: %d ions, %d states, %.1f MB pair data' % (
        n_uc, S, A, system.n_ions, system.N, system.stream_bytes / 1e6), flush=True)
    for label, env in (('default', {}), ('cluster G=8', {'STB_GROUP_CTAS': '8'}),
                       ('grid auto', {'STB_CLUSTER': '0'}),
                       ('grid G=16', {'STB_CLUSTER': '0', 'STB_GROUP_CTAS': '16'}),
                       ('grid G=32', {'STB_CLUSTER': '0', 'STB_GROUP_CTAS': '32'}),
                       ('grid G=64', {'STB_CLUSTER': '0', 'STB_GROUP_CTAS': '64'}),
                       ('grid G=148', {'STB_CLUSTER': '0', 'STB_GROUP_CTAS': '148'})):
        for k in ('STB_CLUSTER', 'STB_GROUP_CTAS'):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            _, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, pop.astype(float), rtol=1e-8,
                                                            atol=1e-15, nsteps=100000, want_y=False)
        except Exception as err:     # pylint: disable=broad-except
            print('   %-12s failed: %s' % (label, err))
            continue
        ms, _ = system.last_timing()
        prof = system.last_profile()
        print('   %-12s %8.1f ms  G=%d groups=%d steps=%d passes=%d  us/pass=%.1f ok=%s' % (
            label, ms, prof['G'], prof['groups'], st[0]['n_steps'], st[0]['n_rhs'] + st[0]['n_jv'],
            1e3 * ms / (st[0]['n_rhs'] + st[0]['n_jv'] + st_p[0]['n_rhs'] + st_p[0]['n_jv']),
            st[0]['status'] == 0), flush=True)
    system.close()
