"""Run-to-run reproducibility of the cluster integrator per cluster size (race detector substitute).
Usage: [STB_LIB=...] python tools/repro_check.py   - prints, per STB_GROUP_CTAS in (1, 2, 4, 8), whether
three repeated solves of 8 systems of the bench lattice are bit-identical and the largest difference."""
import os
import sys
sys.path.insert(0, '.')
import numpy as np
import bench
from simetuc_b200 import lattice, precalculate

cte = bench.make_cte()
powers = bench.powers_list()[::4]
lat = lattice.generate(cte, seed=2000)
ref = None
for G in ('1', '2', '4', '8'):
    os.environ['STB_GROUP_CTAS'] = G
    runs = []
    for rep in range(3):
        system = precalculate.build_system(cte, lat, n_sys=len(powers))
        for k, P in enumerate(powers):
            for exc in cte['excitations']['NIR_980']:
                exc.power_dens = P
            system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
        _, _, pop = system.ground_states()
        y0 = np.tile(pop.astype(float), (len(powers), 1))
        y_p, _, st_p = system.integrate(np.array([0.0, 1e-8]), y0, pulse=True, rtol=1e-8, atol=1e-15)
        t = np.logspace(-8, -2, 60)
        y, avg, st = system.integrate(t, y_p[:, -1], pulse=False, rtol=1e-8, atol=1e-15, want_avg=True)
        system.close()
        runs.append((y_p, y, [(s['n_steps'], s['n_rhs'], s['n_jv'], s['status']) for s in st_p + st]))
    same = all(np.array_equal(runs[0][1], r[1]) and np.array_equal(runs[0][0], r[0]) for r in runs[1:])
    dmax = max(float(np.max(np.abs(runs[0][1] - r[1]))) for r in runs[1:])
    dp = max(float(np.max(np.abs(runs[0][0] - r[0]))) for r in runs[1:])
    if ref is None:
        ref = runs[0][1]
    vs_ref = float(np.max(np.abs(runs[0][1] - ref) / (1e-4 * np.abs(ref) + 1e-12)))
    print('G=%s identical=%s max|dy| relax %.3e pulse %.3e  parity metric vs G=1 %.3g  steps %s' % (
        G, same, dmax, dp, vs_ref, [r[2][-1][0] for r in runs]))
