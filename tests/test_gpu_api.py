"""GPU tests of the reference-facing Python API (`Simulations`, `odesolver`, `optimize`)."""
import os

import numpy as np
import pytest

from oracle import oracle_np as onp
from simetuc_b200 import configs, lattice, odesolver, optimize, precalculate
from simetuc_b200.lattice import Lattice
from simetuc_b200.simulations import (ConcentrationDependenceSolution, DynamicsSolution,
                                      PowerDependenceSolution, Simulations, SteadyStateSolution)
from tests.helpers import GOLDEN, load_golden, oracle_operands

pytestmark = pytest.mark.gpu


@pytest.fixture()
def lattice_file_2S_2A(tmp_path):
    g = load_golden('sim_2S_2A')
    lat = Lattice.from_arrays(g['ion_type'], g['dist_array'], 2, 7, positions=g['doped_lattice'],
                              name='bNaYF4', S_conc=0.3, A_conc=0.3)
    path = str(tmp_path / 'data_2S_2A.npz')
    lat.save(path)
    return path


def test_sim_dyn_2S_2A(lattice_file_2S_2A):
    """Port of the reference's `test_sim_dyn_2S_2A` (`test_simulations.py:76-91`)."""
    cte = configs.sim_cte()
    sim = Simulations(cte, full_path=lattice_file_2S_2A)
    assert sim.cte['lattice']['S_conc'] == 0.3 and sim.cte['states']['sensitizer_states'] == 2
    solution = sim.simulate_dynamics()
    assert isinstance(solution, DynamicsSolution)
    assert solution.index_A_j == [0, -1, -1, 11]
    assert solution.index_S_i == [-1, 7, 9, -1]
    with np.load(os.path.join(GOLDEN, 'sol_2S_2A.npz')) as z:
        assert np.allclose(z['t_sol'], solution.t_sol)
        assert np.allclose(z['y_sol'], solution.y_sol)
    avg = np.array(solution.list_avg_data)
    ref = onp.avg_populations(solution.y_sol, solution.index_S_i, solution.index_A_j, 2, 7)
    assert np.allclose(avg, ref, rtol=1e-12, atol=0)
    steady = solution.calculate_steady_state()
    assert isinstance(steady, SteadyStateSolution)
    assert np.allclose(steady.y_sol, onp.cumulative_steady_state(solution.t_sol, solution.y_sol))
    assert solution.errors.shape == (9,) and np.all(solution.errors == 0)   # no expData here


def test_setup_microscopic_eqs_tuple(lattice_file_2S_2A):
    """The reference's 12-tuple (`precalculate.py:875-878`) with the reference's dtypes."""
    cte = configs.sim_cte()
    out = precalculate.setup_microscopic_eqs(cte, full_path=lattice_file_2S_2A)
    assert len(out) == 12
    (cte2, pop, iS, iA, abs_m, dec_m, ET, N_idx, jac, cET, cN, cJ) = out
    op = oracle_operands(load_golden('sim_2S_2A'), configs.sim_cte())
    assert pop.dtype == np.uint64 and np.array_equal(pop, op['initial_population'])
    assert iS == op['index_S_i'].tolist() and iA == op['index_A_j'].tolist()
    assert abs(abs_m - op['abs']).max() == 0 and abs(dec_m - op['decay']).max() == 0
    assert np.array_equal(N_idx, op['N_indices']) and np.array_equal(jac, op['jac_indices'])
    assert abs(ET - op['ET']).max() <= 1e-14 * abs(op['ET']).max()
    assert cET.shape == (18, 0) and cN.shape == (0, 3) and cJ.shape == (0, 4)
    assert cte2['ions'] == {'total': 4, 'sensitizers': 2, 'activators': 2}
    assert cte2['states']['energy_states'] == 18
    # RHS / dense Jacobian through the reference-named callbacks
    y = load_golden('sim_2S_2A')['sim_y'][0]
    f = odesolver._rate_eq_pulse(0, y, abs_m, dec_m, ET, N_idx, cET, cN)
    ref = onp.rate_eq_pulse(0, y, op['abs'], op['decay'], op['ET'], op['N_indices'], *op['coop'][:2])
    assert np.all(np.abs(f - ref) <= 1e-12 * np.abs(ref).max())
    J = odesolver._jac_rate_eq_pulse(0, y, abs_m, dec_m, ET, jac, cET, cJ)
    Jref = onp.jac_rate_eq_pulse(0, y, op['abs'], op['decay'], op['ET'], op['jac_indices'],
                                 op['coop'][0], op['coop'][2])
    assert np.all(np.abs(J - Jref) <= 1e-12 * np.abs(Jref).max())


def test_solver_mock_seam(lattice_file_2S_2A, monkeypatch):
    """The reference's orchestration tests patch `odesolver._solve_ode` and count calls
    (`test_simulations.py:135-160`): two calls per dynamics simulation."""
    calls = []

    def fake(t_arr, fun, fargs, jfun, jargs, y0, **kw):
        calls.append((len(t_arr), fun.__name__))
        return np.random.random((len(t_arr), len(np.ravel(y0))))

    monkeypatch.setattr(odesolver, '_solve_ode', fake)
    sim = Simulations(configs.sim_cte(), full_path=lattice_file_2S_2A)
    sol = sim.simulate_dynamics()
    assert calls == [(2, '_rate_eq_pulse'), (1000, '_rate_eq')]
    assert sol.y_sol.shape == (1000, 18)


def test_foreign_operands_are_rejected():
    import scipy.sparse as sp
    m = sp.csr_matrix(np.eye(3))
    with pytest.raises(TypeError, match='no CPU fallback'):
        odesolver.solve_relax(np.array([0.0, 1.0]), np.ones(3), m, m, np.zeros((0, 2), np.uint32),
                              np.zeros((0, 3), np.uint32), sp.csr_matrix((3, 0)),
                              np.zeros((0, 3), np.uint32), np.zeros((0, 4), np.uint32))


def test_power_dependence_ESA_closed_form():
    """Port of `test_sim_power_dep_ESA` (`test_simulations.py:347-407`) to the microscopic
    equations: without energy transfer every ion is the same 3-level system and the CW steady
    state is ES1 = GS*P/(k1+P), ES2 = GS*P^2/((k1+P)*k2) with P = pump_rate*power_dens."""
    cte = configs.esa_cte(N_uc=6, A_conc=5.0, t_pulse=None, rtol=1e-8, atol=1e-15)
    cte['lattice_seed'] = 3
    sim = Simulations(cte)
    powers = np.logspace(1, 6, 6)
    sol = sim.simulate_power_dependence(powers)
    assert isinstance(sol, PowerDependenceSolution) and len(sol) == 6
    k1, k2, sigma = 1e3, 1e6, 1e-3
    for s, power in zip(sol, powers):
        GS, ES1, ES2 = [c[-1] for c in s.list_avg_data]
        P = sigma * power
        assert s.power_dens == power
        assert np.isclose(ES1, GS * P / (k1 + P), rtol=1e-4)
        assert np.isclose(ES2, GS * P ** 2 / ((k1 + P) * k2), rtol=1e-4)
        assert np.allclose(s.steady_state_populations, 5.0 * np.array([GS, ES1, ES2]))


def test_power_dependence_pulsed_matches_single_runs():
    cte = configs.yb_tm_cte(N_uc=6, S_conc=10.0, A_conc=2.0)
    cte['lattice_seed'] = 21
    powers = [1e6, 1e9, 1e12]
    sim = Simulations(cte)
    batch = sim.simulate_power_dependence(powers)
    light = Simulations(configs.yb_tm_cte(N_uc=6, S_conc=10.0, A_conc=2.0))
    light.cte['lattice_seed'] = 21
    no_y = light.simulate_power_dependence(powers, keep_y=False)
    for k, P in enumerate(powers):
        one = configs.yb_tm_cte(N_uc=6, S_conc=10.0, A_conc=2.0, power_dens=P)
        one['lattice_seed'] = 21
        single = Simulations(one).simulate_steady_state()
        assert np.array_equal(single.y_sol, batch[k].y_sol)
        assert np.allclose(batch[k].steady_state_populations, no_y[k].steady_state_populations,
                           rtol=1e-10, atol=1e-300)
    pops = batch.populations
    assert pops.shape == (3, 11) and np.all(np.diff(pops[:, 1]) > 0)


def test_concentration_dependence_and_sampling():
    cte = configs.yb_tm_cte(N_uc=5, S_conc=5.0, A_conc=1.0)
    cte['lattice_seed'] = 2
    cte['simulation_params']['N_steps'] = 100
    sim = Simulations(cte)
    concs = [(5.0, 2.0), (20.0, 2.0)]
    sol = sim.simulate_concentration_dependence(concs, N_uc_list=[6, 5], dynamics=True)
    assert isinstance(sol, ConcentrationDependenceSolution) and len(sol) == 2
    assert [tuple(c) for c in sol.concentrations] == concs
    assert sol[0].y_sol.shape[0] == 100 and sol[0].y_sol.shape[1] != sol[1].y_sol.shape[1]
    for s in sol:
        ion_sums = s.y_sol[-1].sum() / (s.cte['ions']['total'])
        assert np.isclose(ion_sums, 1.0, atol=1e-9)          # populations conserved per ion
    avg = sim.sample_simulation(sim.simulate_dynamics, N_samples=2)
    assert isinstance(avg, DynamicsSolution) and len(avg.list_avg_data_ofs) == 11


def test_objective_batch_equals_sequential_objective(tmp_path, monkeypatch):
    """`optimize.optim_fun` one candidate at a time == `objective_batch` on all candidates."""
    cte = configs.yb_tm_cte(N_uc=5, S_conc=20.0, A_conc=2.0)
    cte['lattice_seed'] = 4
    cte['simulation_params']['N_steps'] = 200
    # a synthetic "experimental" decay of Tm 3H4 so that the objective is not identically zero
    exp_dir = tmp_path / 'expData' / 'bNaYF4'
    exp_dir.mkdir(parents=True)
    t = np.linspace(0, 0.02, 400)
    curve = np.exp(-t / 1.5e-3) - np.exp(-t / 2e-4) + 0.01
    np.savetxt(exp_dir / 'decay_Tm_3H4_exc_NIR_980_20.0Yb_2.0Tm.txt', np.column_stack((t, curve)))
    monkeypatch.chdir(tmp_path)
    # the cluster integrator picks its preconditioner (quads / 16-ion groups) per launch from the iteration
    # counts, a batch launch for all its systems at once: pinned here, so that a batch and its systems one
    # at a time take the same path and agree to round-off (otherwise: within the integration tolerance)
    monkeypatch.setenv('STB_PCG', '4')
    sim = Simulations(cte)
    base = [p.strength for p in cte['optimization']['processes']]
    rng = np.random.default_rng(5000)
    cands = [list(np.array(base) * 10.0 ** rng.uniform(-1, 1, len(base))) for _ in range(3)]
    errors, sols = optimize.objective_batch(sim, cands)
    assert errors.shape == (3, 11) and np.any(errors > 0)
    for cand, err in zip(cands, errors):
        seq = optimize.optim_fun(sim.simulate_dynamics, cand, sim)
        assert np.allclose(seq, err, rtol=1e-9, atol=1e-12)


def test_objective_on_device_matches_host_objective(tmp_path, monkeypatch):
    """`objective_batch(on_device=True)` (k_objective: background, interpolation onto the experimental
    time base, per-state deviation on the GPU; only [n, sS+sA] numbers come back) against the host
    objective, whose pieces are pinned against the unmodified reference in
    tests/test_host.py::test_objective_pieces_match_the_reference.  Curves: the four decay curves
    shipped for the config_file.cfg system (stored in tests/golden/expdata_objective.npz), renamed
    to this lattice's concentrations."""
    import os
    from tests.helpers import GOLDEN
    with np.load(os.path.join(GOLDEN, 'expdata_objective.npz')) as z:
        gold = {k: z[k] for k in z.files}
    cte = configs.yb_tm_cte(N_uc=6, S_conc=20.0, A_conc=2.0)
    cte['lattice_seed'] = 6
    cte['simulation_params']['N_steps'] = 300
    exp_dir = tmp_path / 'expData' / 'bNaYF4'
    exp_dir.mkdir(parents=True)
    for state in (str(s) for s in gold['states']):
        np.savetxt(exp_dir / ('decay_%s_exc_NIR_980_20.0Yb_2.0Tm.txt' % state), gold['raw_' + state],
                   fmt='%.18e', delimiter='\t')
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('STB_PCG', '4')      # same preconditioner in both calls: see the test above
    sim = Simulations(cte)
    base = [p.strength for p in cte['optimization']['processes']]
    rng = np.random.default_rng(77)
    cands = [list(np.array(base) * 10.0 ** rng.uniform(-1, 1, len(base))) for _ in range(5)]
    host, sols = optimize.objective_batch(sim, cands)
    dev, none = optimize.objective_batch(sim, cands, on_device=True)
    assert none is None and dev.shape == host.shape == (5, 11)
    assert np.count_nonzero(host[0]) == 4                       # the four states with a curve
    assert np.allclose(dev, host, rtol=1e-10, atol=0.0)


def test_concurrent_handles_match_sequential():
    """Independent lattices solved from several host threads at once (what
    `simulate_concentration_dependence` does with `streams_per_device`) give bit-identical
    results to solving them one after the other: handles share nothing but the memory pool."""
    import threading
    from simetuc_b200.engine import RateSystem

    cte = configs.yb_tm_cte(N_uc=6, S_conc=10.0, A_conc=2.0, power_dens=1e8, t_pulse=1e-8)
    t = np.logspace(-8, -4, 30)

    def solve(seed):
        lat = lattice.generate(cte, seed=seed)
        system = RateSystem(lat, cte['energy_transfer'], device=0)
        try:
            system.set_batch(1)
            system.set_decay_blocks(*precalculate.decay_blocks(cte))
            system.set_abs_blocks(*precalculate.absorption_blocks(cte))
            system.set_strengths([p.strength for p in cte['energy_transfer'].values() if len(p.type) == 2])
            _, _, pop = system.ground_states()
            y, _, st = system.integrate(t, pop.astype(float)[None, :], pulse=True, rtol=1e-8, atol=1e-15)
            assert st[0]['status'] == 0
            return y[0]
        finally:
            system.close()

    seeds = list(range(7000, 7006))
    sequential = [solve(s) for s in seeds]
    parallel = [None] * len(seeds)
    errors = []

    def worker(k):
        try:
            parallel[k] = solve(seeds[k])
        except BaseException as err:      # pylint: disable=broad-except
            errors.append(err)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(seeds))]
    [th.start() for th in threads]
    [th.join() for th in threads]
    assert not errors, errors
    for a, b in zip(sequential, parallel):
        assert np.array_equal(a, b)


def _oracle_average(cte):
    """Oracle operands of the average equations: the two-ion all-ones lattice with the processes at
    their average strengths, interaction columns that cancel identically removed
    (`precalculate.py:971-1021`)."""
    avg_cte = precalculate.average_view(cte)
    n_S, n_A = int(float(cte['lattice']['S_conc']) != 0), int(float(cte['lattice']['A_conc']) != 0)
    ion_type = np.array([0] * n_S + [1] * n_A)
    g = {'ion_type': ion_type, 'dist_array': np.ones((len(ion_type),) * 2),
         'info': np.array([cte['states']['sensitizer_states'], cte['states']['activator_states']])}
    op = oracle_operands(g, avg_cte)
    dense = op['ET'].toarray()
    keep = np.array([not np.allclose(dense[:, k], 0) for k in range(dense.shape[1])], dtype=bool)
    return op, dense[:, keep], op['N_indices'][keep]


def test_average_rate_equations_match_oracle():
    """`average=True` (`simulations.py:1114-1119`, `precalculate.py:884-1051`): operands equal to
    the oracle's, dynamics within the trajectory tolerance of scipy LSODA on the same equations."""
    cte = configs.yb_tm_cte(N_uc=8, S_conc=10.0, A_conc=1.0, power_dens=1e8, t_pulse=1e-8,
                            rtol=1e-8, atol=1e-15, N_steps=300)
    for k, p in enumerate(cte['energy_transfer'].values()):
        p._strength_avg = p.strength * (0.5 + 0.1 * k) * 1e-3      # average strengths differ from microscopic ones
    op, ET_ref, N_ref = _oracle_average(cte)
    out = precalculate.setup_average_eqs(cte)
    (_, pop, iS, iA, abs_m, dec_m, ET, N_idx, jac_idx, cM, cN, cJ) = out
    assert np.array_equal(pop, op['initial_population']) and list(iS) == list(np.ravel(op['index_S_i']))
    assert np.array_equal(N_idx, N_ref) and np.array_equal(jac_idx, onp.jac_indices(N_ref))
    assert np.allclose(ET.toarray(), ET_ref, rtol=1e-14, atol=0)
    assert np.allclose(abs_m.toarray(), op['abs'].toarray(), rtol=1e-15, atol=0)
    assert np.allclose(dec_m.toarray(), op['decay'].toarray(), rtol=1e-15, atol=0)
    assert cM.shape == (11, 0) and cN.shape == (0, 3) and cJ.shape == (0, 4)

    sim = Simulations(cte)
    sol = sim.simulate_dynamics(average=True)
    assert sol.average and sol.y_sol.shape == (300, 11)
    t_pulse, t_sol = onp.time_grids(1e-8, onp.t_simulation(cte['decay']), 300)
    kw = dict(rtol=1e-8, atol=1e-15, nsteps=100000)
    yp = onp.solve_pulse(t_pulse, op['initial_population'].astype(float), op['abs'], op['decay'],
                         op['ET'], op['N_indices'], op['jac_indices'], **kw)
    y_ref = onp.solve_relax(t_sol, yp[-1], op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
    err = np.abs(sol.y_sol - y_ref) / (1e-4 * np.abs(y_ref) + 1e-12)
    assert err.max() <= 1.0
    # sweeps accept the flag too and agree with single runs
    pd = sim.simulate_power_dependence([1e6, 1e8], average=True)
    single = Simulations(cte).simulate_steady_state(average=True)
    assert np.allclose(pd[1].steady_state_populations, single.steady_state_populations, rtol=1e-9)


def test_cooperative_process_through_the_reference_api(lattice_file_2S_2A):
    """A settings tree with a cooperative S-S-A process (`conftest.py:26-117` has `coop1`) through
    `setup_microscopic_eqs` and the reference-named solver entry points: the cooperative operands
    are the oracle's (`precalculate.py:427-699`), the callbacks and `solve_pulse` use them."""
    cte = configs.test_cte()
    cte['lattice']['d_max_coop'] = np.inf
    cte['energy_transfer']['coop1'].strength = 1e15
    out = precalculate.setup_microscopic_eqs(cte, full_path=lattice_file_2S_2A)
    (_, pop, iS, iA, abs_m, dec_m, ET, N_idx, jac, cET, cN, cJ) = out
    op = oracle_operands(load_golden('sim_2S_2A'), cte, coop=True)
    cM_ref, cN_ref, cJ_ref = op['coop']
    assert cN_ref.shape[0] == 2 and np.array_equal(cN, cN_ref) and np.array_equal(cJ, cJ_ref)
    assert np.allclose(cET.toarray(), cM_ref.toarray(), rtol=1e-14, atol=0)
    y = load_golden('sim_2S_2A')['sim_y'][0]
    f = odesolver._rate_eq_pulse(0, y, abs_m, dec_m, ET, N_idx, cET, cN)
    ref = onp.rate_eq_pulse(0, y, op['abs'], op['decay'], op['ET'], op['N_indices'], cM_ref, cN_ref)
    assert np.all(np.abs(f - ref) <= 1e-12 * np.abs(ref).max())
    J = odesolver._jac_rate_eq_pulse(0, y, abs_m, dec_m, ET, jac, cET, cJ)
    Jref = onp.jac_rate_eq_pulse(0, y, op['abs'], op['decay'], op['ET'], op['jac_indices'], cM_ref, cJ_ref)
    assert np.all(np.abs(J - Jref) <= 1e-12 * np.abs(Jref).max())
    t = np.linspace(0, 1e-4, 20)
    kw = dict(rtol=1e-8, atol=1e-15)
    y_dev = odesolver.solve_pulse(t, pop, abs_m, dec_m, ET, N_idx, jac, cET, cN, cJ, nsteps=100000, **kw)
    y_ref = onp.solve_pulse(t, pop.astype(float), op['abs'], op['decay'], op['ET'], op['N_indices'],
                            op['jac_indices'], coop=(cM_ref, cN_ref, cJ_ref), nsteps=100000, **kw)
    assert np.all(np.abs(y_dev - y_ref) <= 1e-4 * np.abs(y_ref) + 1e-12)
    # operands of another system are refused, not silently solved without the term
    with pytest.raises(TypeError):
        odesolver._rate_eq_pulse(0, y, abs_m, dec_m, ET, N_idx, cET, np.zeros((0, 3), np.uint32))


def test_operand_sets_keep_their_own_parameters():
    """Two `setup_microscopic_eqs` calls on one lattice share a device handle; each operand set must
    still integrate with the parameters it was created from (`precalculate.ParamSet`)."""
    from simetuc_b200 import odesolver
    cte1 = configs.yb_tm_cte(N_uc=5, S_conc=20.0, A_conc=2.0, power_dens=1e6)
    cte1['lattice_seed'] = 11
    cte2 = configs.yb_tm_cte(N_uc=5, S_conc=20.0, A_conc=2.0, power_dens=1e9)
    cte2['lattice_seed'] = 11
    t = np.linspace(0, 1e-6, 5)

    def pulse(op):
        return odesolver.solve_pulse(t, op[1].astype(float), *op[4:12], rtol=1e-8, atol=1e-15)

    op1 = precalculate.setup_microscopic_eqs(cte1)
    y1_alone = pulse(op1)
    op2 = precalculate.setup_microscopic_eqs(cte2)          # same lattice: same device handle
    assert op1[6].stb[0] is op2[6].stb[0]
    y2 = pulse(op2)
    y1_again = pulse(op1)                                    # must not pick up cte2's power
    assert np.array_equal(y1_again, y1_alone)
    assert np.abs(y2 - y1_alone).max() > 1e-6


def test_batch_reports_failed_solves(caplog):
    """A solve that cannot reach the end inside a batched sweep is reported like the reference
    reports it (`odesolver.py:129-134`), and a relaxation whose pulse failed is not returned as data
    (ADVICE r1)."""
    import logging
    from simetuc_b200.engine import RateSystem
    cte = configs.yb_tm_cte(N_uc=5, S_conc=20.0, A_conc=2.0, power_dens=1e7)
    cte['lattice_seed'] = 12
    sim = Simulations(cte)
    lat = precalculate.get_lattice(cte)
    orig = RateSystem.integrate_dynamics

    def starved(self, t_pulse, t_sol, y0, **kw):
        kw['nsteps'] = 2                                     # far too few steps between two output times
        return orig(self, t_pulse, t_sol, y0, **kw)

    RateSystem.integrate_dynamics = starved
    try:
        with caplog.at_level(logging.WARNING):
            sols, _, _, stats = sim._solve_batch(lat, [cte], 0, dynamics=True, keep_y=False)
    finally:
        RateSystem.integrate_dynamics = orig
    assert stats[0]['pulse']['status'] == 1 or stats[0]['relax']['status'] == 1
    assert any('increase "nsteps"' in r.getMessage() for r in caplog.records)
    if stats[0]['pulse']['status'] != 0:
        assert np.all(np.isnan(sols[0][2]))


def test_batch_rejects_differing_cooperative_strengths():
    cte = configs.test_cte()
    cte['lattice'] = dict(cte['lattice'])
    cte['lattice'].update(N_uc=4, S_conc=20.0, A_conc=5.0)
    cte['lattice_seed'] = 13
    import copy
    other = copy.deepcopy(cte)
    other['energy_transfer']['coop1'].strength = 1e10
    sim = Simulations(cte)
    lat = precalculate.get_lattice(cte)
    with pytest.raises(ValueError, match='cooperative'):
        sim._solve_batch(lat, [cte, other], 0, dynamics=True)


@pytest.mark.parametrize('env', [{'STB_CL_OPTS': '0'}, {'STB_CL_OPTS': '3'}, {'STB_CL_V': '2'},
                                 {'STB_QUADS': '0'}, {'STB_CLUSTER': '0'}])
def test_integrator_variants_agree(monkeypatch, env):
    """Every option set of the integrators (no Newton start / no lagged inverses, no quad blocks, the
    ion-major cluster kernel, the grid integrator) is another path to the same solution."""
    import bench
    cte = bench.make_cte(1e9)
    t_pulse, t_sol = bench.t_grids(cte)
    res = []
    for variant in ({}, env):
        for k, v in variant.items():
            monkeypatch.setenv(k, v)
        lat = lattice.generate(cte, seed=2000)
        system = precalculate.build_system(cte, lat, n_sys=1)
        _, _, pop = system.ground_states()
        _, y, avg, st_p, st = system.integrate_dynamics(t_pulse, t_sol[::10], pop.astype(float), rtol=1e-8,
                                                        atol=1e-15, want_y=True)
        assert st_p[0]['status'] == 0 and st[0]['status'] == 0
        res.append(y[0])
        system.close()
    assert np.max(np.abs(res[1] - res[0]) / (1e-4 * np.abs(res[0]) + 1e-12)) <= 1.0
