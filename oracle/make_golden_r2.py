"""Golden trajectories for the parity tests that are too slow to solve live in the test suite
(scipy LSODA with the dense Jacobian needs 10 s - 3 min per system at these sizes).  TEST
INFRASTRUCTURE: the solver is the oracle (`oracle/oracle_np.py`: the reference's `_rate_eq*` /
`_jac_rate_eq*` restated and pinned, driven by scipy LSODA as `odesolver.py:85-137` drives `ode`)
at rtol 1e-8 / atol 1e-15.

    python -m oracle.make_golden_r2 [case ...]      # build container; writes tests/golden/traj_<case>.npz

Cases:
  fix240   the reference's own benchmark lattice (`test/test_benchmark/data_240S_108A.hdf5`,
           240 Yb + 108 Tm, 1236 states) with the parameters of `test/conftest.py:288-312`:
           pulse 10 ns + 200-point relaxation
  c1p      C1 at parity scale: config_file.cfg shape, N_uc=12, 25 % Yb, 0.3 % Tm, NIR_980, 1e7 W/cm2,
           pulse + 1000-point relaxation
  c3a/c3b  two corners of the C3 concentration grid (10 % / 0.1 % and 50 % / 2 %), N_uc=10, dynamics
  c2all    all 32 powers of the C2 sweep on the bench lattice (seed 2000): emission curves only
  coop     cooperative (S-S-A) energy transfer on a 277-ion lattice, d_max_coop 8 A: pulse + 100-point
           relaxation, and row 40 of the same relaxation without the cooperative process

Stored per case: the lattice arrays fed to the oracle (ion types, positions; the test regenerates the
lattice with the same seed and checks they agree), the state after the pulse, the full state at 12
output times and the emission curves (ion averages per state, `simulations.py:100-128`).
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

RTOL, ATOL = 1e-8, 1e-15


def rows_of(n):
    return np.unique(np.round(np.array([1, 0.05 * n, 0.15 * n, 0.3 * n, 0.45 * n, 0.6 * n, 0.7 * n, 0.8 * n,
                                        0.85 * n, 0.9 * n, 0.95 * n, n - 1])).astype(int))


def _solve(g, cte, n_steps, t_pulse_w=1e-8):
    from oracle import oracle_np as onp
    from tests.helpers import oracle_operands
    sS, sA = int(g['info'][0]), int(g['info'][1])
    op = oracle_operands(g, cte)
    tf = onp.t_simulation(cte['decay'])
    t_pulse, t_sol = onp.time_grids(t_pulse_w, tf, n_steps)
    kw = dict(rtol=RTOL, atol=ATOL, nsteps=100000)
    y_p = onp.solve_pulse(t_pulse, op['initial_population'].astype(float), op['abs'], op['decay'],
                          op['ET'], op['N_indices'], op['jac_indices'], **kw)
    y = onp.solve_relax(t_sol, y_p[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    avg = onp.avg_populations(y, op['index_S_i'], op['index_A_j'], sS, sA).T
    rows = rows_of(n_steps)
    return dict(t_sol=t_sol, y_pulse_end=y_p[-1], rows=rows, y_rows=y[rows], avg=avg,
                ion_type=np.asarray(g['ion_type']))


def case_fix240():
    from simetuc_b200 import configs
    from tests.helpers import load_golden
    g = load_golden('240S_108A')
    return _solve(g, configs.sim_cte(), 200)


def _ybtm(N_uc, S, A, seed, n_steps, power=1e7):
    from simetuc_b200 import configs, lattice
    cte = configs.yb_tm_cte(N_uc=N_uc, S_conc=S, A_conc=A, power_dens=power, t_pulse=1e-8,
                            rtol=RTOL, atol=ATOL, N_steps=n_steps)
    lat = lattice.generate_arrays(cte, seed=seed)
    g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'], 'info': np.array([2, 9])}
    out = _solve(g, cte, n_steps)
    out['params'] = np.array([N_uc, S, A, seed, n_steps, power], dtype=np.float64)
    return out


def case_c1p():
    return _ybtm(12, 25.0, 0.3, 1012, 1000)


def case_c3a():
    return _ybtm(10, 10.0, 0.1, 3001, 1000)


def case_c3b():
    return _ybtm(10, 50.0, 2.0, 3002, 1000)


def coop_case():
    """Settings + lattice of the cooperative-ET trajectory case (shared with the test)."""
    from simetuc_b200 import configs, lattice
    cte = configs.test_cte()
    cte['lattice'] = dict(cte['lattice'])
    lat = lattice.generate_arrays(configs.yb_tm_cte(N_uc=10, S_conc=15.0, A_conc=3.0), seed=4343)
    g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'], 'info': np.array([2, 7]),
         'doped_lattice': lat['doped_lattice']}
    cte['lattice']['d_max_coop'] = 8.0
    cte['energy_transfer']['coop1'].strength = 5e11
    for exc in cte['excitations']['Vis_473']:
        exc.active = False
    for exc in cte['excitations']['NIR_980']:
        exc.active = True
        exc.power_dens = 1e9
    return cte, g


def case_coop():
    from oracle import oracle_np as onp
    from tests.helpers import oracle_operands
    cte, g = coop_case()
    op = oracle_operands(g, cte, coop=True)
    t_pulse, t_sol = onp.time_grids(1e-8, onp.t_simulation(cte['decay']), 100)
    y0 = op['initial_population'].astype(float)
    kw = dict(rtol=RTOL, atol=ATOL, nsteps=100000)
    yp = onp.solve_pulse(t_pulse, y0, op['abs'], op['decay'], op['ET'], op['N_indices'], op['jac_indices'],
                         coop=op['coop'], **kw)
    ys = onp.solve_relax(t_sol, yp[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'],
                         coop=op['coop'], **kw)
    y_nc = onp.solve_relax(t_sol, yp[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    return dict(t_sol=t_sol, y_pulse_end=yp[-1], y=ys, y_no_coop_row40=y_nc[40], ion_type=np.asarray(g['ion_type']),
                n_coop=np.int64(op['coop'][1].shape[0]))


def _c2_one(k):
    import warnings
    warnings.simplefilter('ignore')
    import bench
    from oracle import oracle_np as onp
    from simetuc_b200 import lattice
    from tests.helpers import oracle_operands
    power = bench.powers_list()[k]
    cte = bench.make_cte(power)
    lat = lattice.generate_arrays(cte, seed=2000)
    g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'], 'info': np.array([2, 9])}
    op = oracle_operands(g, cte)
    t_pulse, t_sol = bench.t_grids(cte)
    kw = dict(rtol=RTOL, atol=ATOL, nsteps=100000)
    y_p = onp.solve_pulse(t_pulse, op['initial_population'].astype(float), op['abs'], op['decay'],
                          op['ET'], op['N_indices'], op['jac_indices'], **kw)
    y = onp.solve_relax(t_sol, y_p[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    avg = onp.avg_populations(y, op['index_S_i'], op['index_A_j'], 2, 9).T
    return power, avg, lat['ion_type'], t_sol


def case_c2all(pool):
    res = pool.map(_c2_one, range(32))
    avg = np.stack([r[1] for r in res])                  # [32, 1000, 11]
    keep = np.unique(np.concatenate([np.arange(0, 1000, 4), [999]]))
    return dict(powers=np.array([r[0] for r in res]), t_rows=keep, avg=avg[:, keep, :],
                ion_type=res[0][2], t_sol=res[0][3])


def _run_case(name):
    import warnings
    warnings.simplefilter('ignore')
    t0 = time.time()
    out = globals()['case_' + name]()
    return name, out, time.time() - t0


def main():
    cases = sys.argv[1:] or ['fix240', 'c1p', 'c3a', 'c3b', 'coop', 'c2all']
    single = [c for c in cases if c != 'c2all']
    with Pool(8) as pool:
        pending = [pool.apply_async(_run_case, (c,)) for c in single]
        for p in pending:
            name, out, dt = p.get()
            path = os.path.join(ROOT, 'tests', 'golden', 'traj_%s.npz' % name)
            np.savez_compressed(path, **out)
            print('wrote %s (%d bytes, %.0f s of LSODA)' % (path, os.path.getsize(path), dt), flush=True)
        if 'c2all' in cases:
            t0 = time.time()
            out = case_c2all(pool)
            path = os.path.join(ROOT, 'tests', 'golden', 'traj_c2all.npz')
            np.savez_compressed(path, **out)
            print('wrote %s (%d bytes, %.0f s)' % (path, os.path.getsize(path), time.time() - t0), flush=True)


if __name__ == '__main__':
    main()
