"""Sweep of the GMRES tolerance of the cluster integrator on the C2 batch: time, iteration counts and the
parity metric against the committed oracle curves of all 32 powers (DEVELOPMENT TOOL)."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
import bench                                     # noqa: E402
cte, lat, system = bench.build_sweep_system(0, list(range(32)))
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
y0 = np.tile(pop.astype(float), (32, 1))
with np.load('tests/golden/traj_c2all.npz') as z:
    gold = {k: z[k] for k in z.files}
with np.load('tests/golden/c2_power_sweep.npz') as z:
    gold3 = {k: z[k] for k in z.files}
_, _, sys3 = bench.build_sweep_system(0, [int(k) for k in gold3['power_idx']])
for lin in [float(v) for v in os.environ.get('LINS', '2e-3,5e-2,7e-2,1e-1,2e-1').split(',')]:
    os.environ['STB_LIN_RMS'] = str(lin)
    best = (1e9, 1e9)
    for rep in range(2):
        _, _, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol, y0, rtol=1e-8, atol=1e-15, want_y=False)
        ms = system.last_timing_dynamics()
        if sum(ms) < sum(best):
            best = ms
    ref = gold['avg']
    got = avg[:, gold['t_rows'], :]
    big = ref > 1e-9
    worst = float(np.max(np.abs(got - ref)[big] / ref[big]))
    _, y3, _, _, _ = sys3.integrate_dynamics(t_pulse, t_sol, y0[:3], rtol=1e-8, atol=1e-15, want_y=True)
    rows = gold3['rows']
    ymetric = max(float(np.max(np.abs(y3[k][rows] - gold3['y_rows'][k]) / (1e-4 * np.abs(gold3['y_rows'][k]) + 1e-12)))
                  for k in range(3))
    print('lin_rms %.0e: pulse %.2f relax %.2f ms -> %.0f solves/s | steps %d rhs %d jv %d rej %d nf %d | worst rel err of the curves %.2e (limit 1e-4), state parity metric %.3f (limit 1)' % (
        lin, best[0], best[1], 32e3 / sum(best), sum(s['n_steps'] for s in st), sum(s['n_rhs'] for s in st),
        sum(s['n_jv'] for s in st), sum(s['n_rejected'] for s in st), sum(s['n_newton_fail'] for s in st), worst, ymetric), flush=True)
