"""C2 batch (32 powers) through the cluster integrator for several STB_CL_OPTS settings: kernel
times, step / pass counts, per-phase cycles (STB_VERBOSE) and the parity metric against the
committed oracle trajectories of powers 0 / 16 / 31.  DEVELOPMENT TOOL."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import lattice, precalculate   # noqa: E402
import bench                                     # noqa: E402

opts_list = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]
cte = bench.make_cte()
powers = bench.powers_list()
n = len(powers)
lat = lattice.generate(cte, seed=2000)
system = precalculate.build_system(cte, lat, n_sys=n)
for k, P in enumerate(powers):
    for exc in cte['excitations']['NIR_980']:
        exc.power_dens = P
    system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
_, _, pop = system.ground_states()
y0 = np.tile(pop.astype(float), (n, 1))
t_pulse, t_sol = bench.t_grids(cte)
with np.load('tests/golden/c2_power_sweep.npz') as z:
    gold = {k: z[k] for k in z.files}
for opts in opts_list:
    os.environ['STB_CL_OPTS'] = str(opts)
    os.environ.pop('STB_VERBOSE', None)
    best = None
    for rep in range(3):
        if rep == 2:
            os.environ['STB_VERBOSE'] = '1'
        yp, _, stp = system.integrate(t_pulse, y0, pulse=True, rtol=1e-8, atol=1e-15)
        ms_p = system.last_timing()[0]
        ys, avg, st = system.integrate(t_sol, yp[:, -1], pulse=False, rtol=1e-8, atol=1e-15, want_y=False, want_avg=True)
        ms_r = system.last_timing()[0]
        best = (ms_p, ms_r) if best is None else (min(best[0], ms_p), min(best[1], ms_r))
    worst = 0.0
    for slot, k in enumerate(gold['power_idx']):
        ref = gold['avg'][slot]
        big = ref > 1e-9
        worst = max(worst, float(np.max(np.abs(avg[k] - ref)[big] / ref[big])))
    print('OPTS %d: pulse %.2f ms relax %.2f ms -> %.0f solves/s | steps max %d sum %d | rhs %d jv %d | rej %d nf %d | '
          'status ok %s | avg rel err vs oracle %.2e' % (
              opts, best[0], best[1], 32e3 / (best[0] + best[1]), max(s['n_steps'] for s in st),
              sum(s['n_steps'] for s in st), sum(s['n_rhs'] for s in st), sum(s['n_jv'] for s in st),
              sum(s['n_rejected'] for s in st), sum(s['n_newton_fail'] for s in st),
              all(s['status'] == 0 for s in st + stp), worst), flush=True)
    print('   steps per system:', [s['n_steps'] for s in st], flush=True)
