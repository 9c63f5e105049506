"""GPU parity tests of BOTH integrators on the BASELINE.json configurations (run with `-m gpu`).

Every trajectory case runs through the cluster integrator (`k_integrate_cl_t`, the default for
small dense lattices) AND through the persistent grid integrator (`k_integrate`: what large
lattices, cutoff lists and several multipolarities use), forced with `STB_CLUSTER=0` and reached
natively by a `d_max` lattice and by a two-multipolarity process set.  References:

  * committed oracle trajectories (`oracle/make_golden_r2.py`: the oracle's restated
    `_rate_eq* / _jac_rate_eq*` driven by scipy LSODA at rtol 1e-8 / atol 1e-15, the protocol of
    `odesolver.py:85-137`), or the oracle solved live where it takes seconds;
  * tolerance: |delta| <= 1e-4 |ref| + 1e-12 on every stored state entry, 1e-4 relative on the
    emission curves above 1e-9 (SURVEY.md 8d parity protocol);
  * at N_uc = 40 (C4: 24 000 ions, 5.9e8 interactions - no reference operand fits in memory):
    oracle RHS / Jv rows of 64 random ions at 1e-12, and agreement of the emission curves under
    tolerance tightening (rtol 1e-8 vs 1e-10).
"""
import os

import numpy as np
import pytest

from oracle import oracle_np as onp
from simetuc_b200 import configs, lattice, precalculate
from simetuc_b200.engine import RateSystem
from tests.helpers import GOLDEN, load_golden, oracle_operands, pair_processes
from tests.test_gpu_parity import make_system, parity_metric

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-8, 1e-15


def _gold(name):
    with np.load(os.path.join(GOLDEN, 'traj_%s.npz' % name)) as z:
        return {k: z[k] for k in z.files}


def _integrators(monkeypatch, which):
    """Select the integrator through the library's environment switches."""
    if which == 'grid':
        monkeypatch.setenv('STB_CLUSTER', '0')
    elif which == 'grid-tma':
        # the grid integrator as the large lattices run it: TMA-fed passes, group blocks with the balanced
        # pair pass and the neighbour hand-over, activator rows on their own CTAs, lagged inverses
        monkeypatch.setenv('STB_CLUSTER', '0')
        monkeypatch.setenv('STB_TMA', '1')
    else:
        monkeypatch.delenv('STB_CLUSTER', raising=False)


def _check_traj(system, gold, n_sys=1, first=0):
    _, _, pop = system.ground_states()
    t_sol = gold['t_sol']
    t_pulse = np.linspace(0.0, t_sol[0], 2)
    y0 = pop.astype(float)
    y_p, _, st_p = system.integrate(t_pulse, y0, pulse=True, rtol=RTOL, atol=ATOL, first=first)
    y, avg, st = system.integrate(t_sol, y_p[0, -1], pulse=False, rtol=RTOL, atol=ATOL, want_avg=True,
                                  first=first)
    assert st_p[0]['status'] == 0 and st[0]['status'] == 0 and st[0]['n_out'] == len(t_sol)
    assert parity_metric(y_p[0, -1], gold['y_pulse_end']) <= 1.0
    assert parity_metric(y[0][gold['rows']], gold['y_rows']) <= 1.0
    ref = gold['avg']
    big = ref > 1e-9
    assert np.all(np.abs(avg[0] - ref)[big] <= 1e-4 * ref[big])
    return st[0]


@pytest.mark.parametrize('which', ['cluster', 'grid'])
def test_reference_benchmark_lattice(monkeypatch, which):
    """240 Yb + 108 Tm (`test/test_benchmark/data_240S_108A.hdf5`), parameters of
    `test/conftest.py:288-312`: pulse + 200-point relaxation against the committed LSODA run."""
    gold = _gold('fix240')
    g = load_golden('240S_108A')
    system = make_system(g, configs.sim_cte())
    _integrators(monkeypatch, which)
    st = _check_traj(system, gold)
    assert st['n_steps'] > 50
    system.close()


def _ybtm_system(gold):
    N_uc, S, A, seed, n_steps, power = gold['params']
    cte = configs.yb_tm_cte(N_uc=int(N_uc), S_conc=float(S), A_conc=float(A), power_dens=float(power),
                            t_pulse=1e-8, rtol=RTOL, atol=ATOL, N_steps=int(n_steps))
    lat = lattice.generate(cte, seed=int(seed))
    assert np.array_equal(lat.ion_type, gold['ion_type'])        # the lattice the oracle solved
    return precalculate.build_system(cte, lat, n_sys=1)


@pytest.mark.parametrize('which', ['cluster', 'grid', 'grid-tma'])
def test_c1_parity_scale(monkeypatch, which):
    """C1 (config_file.cfg shape: 25 % Yb, 0.3 % Tm, 980 nm pulse) at the survey's parity scale
    N_uc = 12 (~1370 states): pulse + 1000-point relaxation."""
    gold = _gold('c1p')
    system = _ybtm_system(gold)
    _integrators(monkeypatch, which)
    _check_traj(system, gold)
    system.close()


@pytest.mark.parametrize('which', ['cluster', 'grid', 'grid-tma'])
@pytest.mark.parametrize('case', ['c3a', 'c3b'])
def test_c3_grid_corners(monkeypatch, which, case):
    """Two corners of the C3 concentration grid (10 % / 0.1 % and 50 % / 2 %), dynamics."""
    gold = _gold(case)
    system = _ybtm_system(gold)
    _integrators(monkeypatch, which)
    _check_traj(system, gold)
    system.close()


def test_c2_all_32_powers():
    """The bench workload, every power of the sweep: emission curves of the 32 systems of the bench
    lattice (seed 2000) against the oracle's (`traj_c2all.npz`, every 4th output time)."""
    import bench
    gold = _gold('c2all')
    powers = bench.powers_list()
    assert np.allclose(powers, gold['powers'], rtol=1e-15)
    cte = bench.make_cte()
    lat = lattice.generate(cte, seed=2000)
    assert np.array_equal(lat.ion_type, gold['ion_type'])
    system = precalculate.build_system(cte, lat, n_sys=len(powers))
    for k, P in enumerate(powers):
        for exc in cte['excitations']['NIR_980']:
            exc.power_dens = P
        system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=k)
    _, _, pop = system.ground_states()
    t_pulse, t_sol = bench.t_grids(cte)
    assert np.array_equal(t_sol, gold['t_sol'])
    y0 = np.tile(pop.astype(float), (len(powers), 1))
    y_p, _, st_p = system.integrate(t_pulse, y0, pulse=True, rtol=RTOL, atol=ATOL)
    _, avg, st = system.integrate(t_sol, y_p[:, -1], pulse=False, rtol=RTOL, atol=ATOL, want_y=False,
                                  want_avg=True)
    system.close()
    assert all(s['status'] == 0 for s in st_p + st)
    ref = gold['avg']
    got = avg[:, gold['t_rows'], :]
    big = ref > 1e-9
    assert np.all(np.abs(got - ref)[big] <= 1e-4 * ref[big])


def _live_oracle(g, cte, n_steps):
    op = oracle_operands(g, cte)
    tf = onp.t_simulation(cte['decay'])
    t_pulse, t_sol = onp.time_grids(1e-8, tf, n_steps)
    kw = dict(rtol=RTOL, atol=ATOL, nsteps=100000)
    y0 = op['initial_population'].astype(float)
    yp = onp.solve_pulse(t_pulse, y0, op['abs'], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    ys = onp.solve_relax(t_sol, yp[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    return op, t_pulse, t_sol, yp, ys


def test_cutoff_lattice_trajectory():
    """A `d_max` lattice is stored as indexed lists and integrated by the grid integrator natively
    (no environment switch): pulse + 200-point relaxation against the oracle fed the same filtered
    interaction lists."""
    cte = configs.yb_tm_cte(N_uc=7, S_conc=15.0, A_conc=2.0, power_dens=1e7)
    lat = lattice.generate(cte, seed=8, with_distances=True)
    d_max = 12.0
    system = RateSystem(lat, dict(pair_processes(cte)), d_max=d_max)
    system.set_decay_blocks(*precalculate.decay_blocks(cte))
    system.set_abs_blocks(*precalculate.absorption_blocks(cte))
    assert system.bytes_per_pair == 12                       # indexed lists, not dense blocks
    dist = np.where(lat.dist_array <= d_max, lat.dist_array, 0.0)
    g = {'ion_type': lat.ion_type, 'dist_array': dist, 'info': np.array([2, 9])}
    op, t_pulse, t_sol, yp_ref, ys_ref = _live_oracle(g, cte, 200)
    y0 = op['initial_population'].astype(float)
    y_p, _, st_p = system.integrate(t_pulse, y0, pulse=True, rtol=RTOL, atol=ATOL)
    y, _, st = system.integrate(t_sol, y_p[0, -1], pulse=False, rtol=RTOL, atol=ATOL)
    system.close()
    assert st_p[0]['status'] == 0 and st[0]['status'] == 0
    assert parity_metric(y_p[0, -1], yp_ref[-1]) <= 1.0
    assert parity_metric(y[0], ys_ref) <= 1.0


def test_two_multipolarities_trajectory():
    """Processes of different multipolarity (d^-6 and d^-8) keep one rate table per multipolarity;
    such a lattice is integrated by the grid integrator natively."""
    cte = configs.yb_tm_cte(N_uc=7, S_conc=15.0, A_conc=2.0, power_dens=1e7)
    cte['energy_transfer']['ETU2'].mult = 8
    cte['energy_transfer']['CR50'].mult = 8
    lat = lattice.generate(cte, seed=9, with_distances=True)
    system = precalculate.build_system(cte, lat, n_sys=1)
    g = {'ion_type': lat.ion_type, 'dist_array': lat.dist_array, 'info': np.array([2, 9])}
    op, t_pulse, t_sol, yp_ref, ys_ref = _live_oracle(g, cte, 200)
    N_idx, _, w, _ = system.export_interactions()
    assert np.array_equal(N_idx, op['N_indices'])
    y0 = op['initial_population'].astype(float)
    y_p, _, st_p = system.integrate(t_pulse, y0, pulse=True, rtol=RTOL, atol=ATOL)
    y, _, st = system.integrate(t_sol, y_p[0, -1], pulse=False, rtol=RTOL, atol=ATOL)
    system.close()
    assert st_p[0]['status'] == 0 and st[0]['status'] == 0
    assert parity_metric(y_p[0, -1], yp_ref[-1]) <= 1.0
    assert parity_metric(y[0], ys_ref) <= 1.0


def test_c4_large_lattice_rows_and_tolerance_tightening():
    """C4 shape (N_uc = 40, 25 % Yb, 0.3 % Tm: ~24 000 ions, 5.9e8 ordered pairs, 4.7 GB of pair
    data).  (i) RHS and Jv rows of 64 random ions against the oracle's row-sampled operators at
    1e-12; (ii) the grid integrator's emission curves at rtol 1e-8 against rtol 1e-10."""
    cte = configs.yb_tm_cte(N_uc=40, S_conc=25.0, A_conc=0.3, power_dens=1e7)
    lat = lattice.generate(cte, seed=4000)
    system = precalculate.build_system(cte, lat, n_sys=1)
    N = system.N
    rng = np.random.default_rng(40)
    # populations: log-uniform, normalised per ion
    iS, iA, pop = system.ground_states()
    y = 10.0 ** rng.uniform(-20, 0, N)
    gs = np.where(lat.ion_type != 0, iA, iS)
    width = np.where(lat.ion_type != 0, 9, 2)
    owner = np.repeat(np.arange(len(gs)), width)
    y /= np.bincount(owner, weights=y)[owner]
    v = rng.standard_normal(N)
    f = system.rhs(y, pulse=True)
    jv = system.jv(y, v, pulse=True)
    aS, aA = onp.absorption_blocks(2, 9, cte['excitations'])
    dS, dA = onp.decay_blocks(2, 9, cte['decay'])
    ions = np.concatenate([rng.choice(np.flatnonzero(lat.ion_type == 0), 44, replace=False),
                           rng.choice(np.flatnonzero(lat.ion_type != 0), 20, replace=False)])
    idx, f_rows, jv_rows = onp.rate_eq_rows(ions, lat.ion_type, lat.positions, lat.cell, 2, 9,
                                            pair_processes(cte), aS + dS, aA + dA, y, v)
    assert np.all(np.abs(f[idx] - f_rows) <= 1e-12 * np.abs(f_rows).max())
    assert np.all(np.abs(jv[idx] - jv_rows) <= 1e-12 * np.abs(jv_rows).max())
    # (ii) convergence under tolerance tightening
    tf = onp.t_simulation(cte['decay'])
    t_pulse, t_sol = onp.time_grids(1e-8, tf, 200)
    curves = {}
    for rtol in (1e-8, 1e-10):
        y_p, _, st_p = system.integrate(t_pulse, pop.astype(float), pulse=True, rtol=rtol, atol=ATOL)
        _, avg, st = system.integrate(t_sol, y_p[0, -1], pulse=False, rtol=rtol, atol=ATOL,
                                      want_y=False, want_avg=True, nsteps=5000)
        assert st_p[0]['status'] == 0 and st[0]['status'] == 0
        curves[rtol] = avg[0]
    system.close()
    ref = curves[1e-10]
    big = ref > 1e-9
    assert np.all(np.abs(curves[1e-8] - ref)[big] <= 1e-4 * ref[big])


def test_quad_preconditioner_same_solution_fewer_iterations(monkeypatch):
    """Quad lattices (the library stores strongly coupled sensitizers next to each other and
    preconditions them in blocks of four): same interaction lists in the reference's order, same
    trajectory within the integration tolerance, about half the Krylov iterations."""
    import bench
    cte = bench.make_cte()
    res = {}
    for quads in ('1', '0'):
        monkeypatch.setenv('STB_QUADS', quads)
        lat = lattice.generate(cte, seed=2000)
        system = precalculate.build_system(cte, lat, n_sys=1)
        _, _, pop = system.ground_states()
        t_pulse, t_sol = bench.t_grids(cte)
        _, y, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol[::5], pop.astype(float), rtol=RTOL,
                                                        atol=ATOL, want_y=True)
        assert st_p[0]['status'] == 0 and st[0]['status'] == 0
        res[quads] = (system.export_interactions(), y[0], st[0]['n_jv'], st[0]['n_steps'])
        system.close()
    for a, b in zip(res['1'][0], res['0'][0]):
        assert np.array_equal(a, b)                      # reference order, whatever the storage order
    assert parity_metric(res['1'][1], res['0'][1]) <= 1.0
    assert res['1'][2] < 0.75 * res['0'][2]              # Krylov iterations
    assert abs(res['1'][3] - res['0'][3]) <= 0.05 * res['0'][3]


@pytest.mark.parametrize('which', ['cluster', 'grid'])
def test_coop_in_both_integrators(monkeypatch, which):
    """Cooperative (three-ion, S-S-A) energy transfer (`precalculate.py:427-599`, `odesolver.py:40-41,
    56-60`) through the cluster integrator and through the grid integrator (what lattices too large for
    a cluster use: the manual's published dynamics run has 1230 ions and 10 453 cooperative
    interactions) against the committed scipy-LSODA run on the oracle's operators
    (`oracle/make_golden_r2.py coop`; 40 s of CPU)."""
    from oracle.make_golden_r2 import coop_case
    from tests.test_gpu_parity import make_coop_system
    gold = _gold('coop')
    cte, g = coop_case()
    assert np.array_equal(g['ion_type'], gold['ion_type']) and len(g['ion_type']) > 262
    system = make_coop_system(g, cte, d_max_coop=8.0)
    assert system.n_coop == int(gold['n_coop']) > 100
    _integrators(monkeypatch, which)
    if which == 'grid':
        monkeypatch.setenv('STB_GROUP_CTAS', '12')         # a group size only the grid integrator takes
    t_sol = gold['t_sol']
    t_pulse = np.linspace(0.0, t_sol[0], 2)
    _, _, pop = system.ground_states()
    y_p, _, st_p = system.integrate(t_pulse, pop.astype(float), pulse=True, rtol=RTOL, atol=ATOL)
    y, _, st = system.integrate(t_sol, y_p[0, -1], pulse=False, rtol=RTOL, atol=ATOL)
    prof = system.last_profile()
    system.close()
    assert st_p[0]['status'] == 0 and st[0]['status'] == 0
    assert (prof['G'] == 12) == (which == 'grid')          # the integrator that was asked for
    assert parity_metric(y_p[0, -1], gold['y_pulse_end']) <= 1.0
    assert parity_metric(y[0], gold['y']) <= 1.0
    # the cooperative channel matters: without it the state differs far beyond the tolerance
    assert parity_metric(gold['y_no_coop_row40'], gold['y'][40]) > 100.0


@pytest.mark.parametrize('periodic', [True, False])
def test_cell_list_build_equals_the_scan(monkeypatch, periodic):
    """`d_max` lattices given by positions are built with cell lists (O(n x partners)); the pair lists,
    their order and the rates must be bit-identical to the O(n^2) scan (`STB_CELLS=0`), for the
    periodic hexagonal supercell and for an open (nanoparticle-like) lattice."""
    cte = configs.yb_tm_cte(N_uc=14, S_conc=20.0, A_conc=2.0, power_dens=1e7)
    lat = lattice.generate(cte, seed=14)
    if not periodic:
        lat.cell = None
    d_max = 9.0
    et = dict(pair_processes(cte))
    res = {}
    for cells in ('1', '0'):
        monkeypatch.setenv('STB_CELLS', cells)
        system = RateSystem(lat, et, d_max=d_max, geometry='positions')
        assert system.get_option('cell_lists') == float(cells)
        res[cells] = system.export_interactions() + (system.n_pairs,)
        system.close()
    assert res['1'][4] == res['0'][4] > 1000
    for a, b in zip(res['1'][:4], res['0'][:4]):
        assert np.array_equal(a, b)
