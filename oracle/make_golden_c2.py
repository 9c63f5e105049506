"""Golden vectors for the bench workload at FULL size (BASELINE.json configs[1], SURVEY.md 8d "C2"):
`config_file_Yb_Tm.cfg` lattice (N_uc=15, 5 % Yb, 0.5 % Tm, seed 2000 - the lattice `bench.py`
solves on rank 0), pulse 10 ns + 1000-point relaxation at rtol 1e-8 / atol 1e-15, at three of the
32 powers of the sweep (first, middle, last).  The solver is the oracle (`oracle/oracle_np.py`: the
reference's `_rate_eq*` / `_jac_rate_eq*` restated and pinned, driven by scipy LSODA exactly as
`odesolver.py:85-137` drives `ode`), which needs 2-6 minutes of CPU per system: too slow for the
test suite, so its outputs are committed.

    python -m oracle.make_golden_c2          # build container; writes tests/golden/c2_power_sweep.npz

Stored per power: the emission curves (ion averages per state, `simulations.py:100-128`) at all
1000 times, the full state vector at 12 of the output times, the state after the pulse.
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

ROWS = [1, 50, 150, 300, 450, 600, 700, 800, 850, 900, 950, 999]
POWER_IDX = [0, 16, 31]


def solve(k):
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
    kw = dict(rtol=bench.RTOL, atol=bench.ATOL, nsteps=100000)
    y_p = onp.solve_pulse(t_pulse, op['initial_population'].astype(float), op['abs'], op['decay'],
                          op['ET'], op['N_indices'], op['jac_indices'], **kw)
    y = onp.solve_relax(t_sol, y_p[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    avg = onp.avg_populations(y, op['index_S_i'], op['index_A_j'], 2, 9).T
    return dict(power=power, y_pulse_end=y_p[-1], y_rows=y[ROWS], avg=avg, t_sol=t_sol,
                ion_type=lat['ion_type'])


def main():
    with Pool(len(POWER_IDX)) as pool:
        res = pool.map(solve, POWER_IDX)
    out = {'power_idx': np.array(POWER_IDX), 'rows': np.array(ROWS), 't_sol': res[0]['t_sol'],
           'ion_type': res[0]['ion_type'], 'powers': np.array([r['power'] for r in res]),
           'y_pulse_end': np.stack([r['y_pulse_end'] for r in res]),
           'y_rows': np.stack([r['y_rows'] for r in res]),
           'avg': np.stack([r['avg'] for r in res])}
    path = os.path.join(ROOT, 'tests', 'golden', 'c2_power_sweep.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
