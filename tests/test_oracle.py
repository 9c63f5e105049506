"""The oracle (`oracle/oracle_np.py`) against the reference's golden vectors.

`tests/golden/*.npz` were produced by the UNMODIFIED reference functions
(`oracle/make_golden.py`); the trajectory is the reference's own golden file
(`test/test_simulations/y_sol_2S_2A.hdf5`, `test_simulations.py:76-91`)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle_np as onp
from simetuc_b200 import configs
from tests.helpers import GOLDEN, SMALL, load_golden, oracle_operands, pair_processes


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('tag', SMALL)
@pytest.mark.parametrize('pname', ['test', 'sim'])
def test_builders_match_reference(tag, pname):
    g = load_golden(tag)
    cte = getattr(configs, pname + '_cte')()
    op = oracle_operands(g, cte)
    # integers: bit-exact
    assert np.array_equal(op['index_S_i'], g['indices_S_i'])
    assert np.array_equal(op['index_A_j'], g['indices_A_j'])
    assert np.array_equal(op['initial_population'], g['initial_population'])
    assert op['initial_population'].dtype == np.uint64
    for key, lst in zip(('S_k', 'S_l', 'A_k', 'A_l'), op['lists'][:4]):
        ref = g['index_' + key]
        assert len(lst) == ref.shape[0]
        for a, b in zip(lst, ref):
            assert np.array_equal(a, b)
    for key, lst in zip(('S_k', 'S_l', 'A_k', 'A_l'), op['lists'][4:]):
        for a, b in zip(lst, g['dist_' + key]):
            assert np.array_equal(a, b)
    assert np.array_equal(op['N_indices'], g[pname + '_N_indices'])
    assert op['N_indices'].dtype == np.uint32
    assert np.array_equal(op['jac_indices'], g[pname + '_jac_indices'])
    # floats: same formulas in the same order -> bit-exact on one machine
    assert np.array_equal(op['abs'].toarray(), g[pname + '_abs'])
    assert np.array_equal(op['decay'].toarray(), g[pname + '_decay'])
    assert np.array_equal(op['ET'].toarray(), g[pname + '_ET'])


@pytest.mark.parametrize('tag', SMALL + ['240S_108A'])
@pytest.mark.parametrize('pname', ['test', 'sim'])
def test_rhs_and_jv_match_reference(tag, pname):
    g = load_golden(tag)
    cte = getattr(configs, pname + '_cte')()
    op = oracle_operands(g, cte)
    cM, cN, cJ = op['coop']
    for y, v, fp, fr, jv in zip(g[pname + '_y'], g[pname + '_v'], g[pname + '_f_pulse'],
                                g[pname + '_f_relax'], g[pname + '_jv_pulse']):
        f1 = onp.rate_eq_pulse(0.0, y, op['abs'], op['decay'], op['ET'], op['N_indices'], cM, cN)
        f2 = onp.rate_eq(0.0, y, op['decay'], op['ET'], op['N_indices'], cM, cN)
        assert np.array_equal(f1, fp)
        assert np.array_equal(f2, fr)
        lin = op['abs'] + op['decay']
        jv2 = onp.jac_vec(y, v, lin, op['ET'], op['N_indices'])
        scale = np.abs(jv).max() if jv.size else 1.0
        assert np.all(np.abs(jv2 - jv) <= 1e-13 * scale)
        if op['N'] <= 100:
            J = onp.jac_rate_eq_pulse(0.0, y, op['abs'], op['decay'], op['ET'],
                                      op['jac_indices'], cM, cJ)
            assert np.all(np.abs(J @ v - jv) <= 1e-13 * scale)


@pytest.mark.parametrize('pname', ['test', 'sim'])
def test_big_fixture_digests(pname):
    g = load_golden('240S_108A')
    op = oracle_operands(g, getattr(configs, pname + '_cte')())
    ET = op['ET'].tocsr()
    ET.sort_indices()
    assert op['N_indices'].shape[0] == int(g[pname + '_M'][0])
    assert sha(op['N_indices']) == str(g[pname + '_sha_N_indices'])
    assert sha(op['jac_indices']) == str(g[pname + '_sha_jac_indices'])
    assert sha(ET.indices.astype(np.int32)) == str(g[pname + '_sha_ET_indices'])
    assert sha(ET.indptr.astype(np.int32)) == str(g[pname + '_sha_ET_indptr'])
    assert np.isclose(np.abs(ET.data).sum(), g[pname + '_ET_data_sum_abs'][0], rtol=1e-13)
    assert np.allclose(np.abs(op['ET'].tocsc().data[:4000:4]), g[pname + '_ET_colnorm_head'],
                       rtol=1e-15, atol=0)


def test_structural_invariants():
    """Column sums of abs/decay/ET vanish, decay is upper triangular, 4 nnz per ET column
    (`test_setup.py:687-700`)."""
    g = load_golden('240S_108A')
    op = oracle_operands(g, configs.sim_cte())
    for m in (op['abs'], op['decay'], op['ET']):
        assert np.allclose(np.asarray(m.sum(axis=0)), 0, atol=1e-6 * abs(m).max())
    assert (abs(op['decay'] - __import__('scipy.sparse').sparse.triu(op['decay']))).nnz == 0
    assert np.all(np.diff(op['ET'].tocsc().indptr) == 4)


def test_known_answer_2S_2A_pairs():
    """`N_indices` of the hand-checked 2S_2A case (`test_setup.py:522-535`)."""
    g = load_golden('sim_2S_2A')
    op = oracle_operands(g, configs.sim_cte())
    expect = [[5, 11], [5, 14], [3, 7], [3, 9], [8, 9], [8, 0], [8, 11], [10, 7], [10, 0],
              [10, 11], [16, 0], [16, 3], [14, 7], [14, 9]]
    assert op['N_indices'].tolist() == expect
    assert op['ET'].shape == (18, 14) and op['ET'].nnz == 56


def test_golden_trajectory_2S_2A():
    """`test_sim_dyn_2S_2A` (`test_simulations.py:76-91`): pulse 10 ns + 1000-point relaxation at
    rtol 1e-3 / atol 1e-15 must be `np.allclose` to the shipped golden file."""
    g = load_golden('sim_2S_2A')
    cte = configs.sim_cte()
    op = oracle_operands(g, cte)
    with np.load(os.path.join(GOLDEN, 'sol_2S_2A.npz')) as z:
        t_gold, y_gold = z['t_sol'], z['y_sol']
    assert op['index_A_j'].tolist() == [0, -1, -1, 11]
    assert op['index_S_i'].tolist() == [-1, 7, 9, -1]
    tf = onp.t_simulation(cte['decay'])
    t_pulse, t_sol = onp.time_grids(1e-8, tf, 1000)
    assert np.allclose(t_sol - t_sol[0], t_gold)
    for integrator in ('vode', 'lsoda'):
        kw = dict(rtol=1e-3, atol=1e-15, nsteps=5000, integrator=integrator)
        y_p = onp.solve_pulse(t_pulse, op['initial_population'].astype(float), op['abs'],
                              op['decay'], op['ET'], op['N_indices'], op['jac_indices'], **kw)
        y = onp.solve_relax(t_sol, y_p[-1], op['decay'], op['ET'], op['N_indices'],
                            op['jac_indices'], **kw)
        assert np.allclose(y, y_gold), integrator


def test_results_helpers():
    g = load_golden('sim_2S_2A')
    op = oracle_operands(g, configs.sim_cte())
    rng = np.random.default_rng(3)
    y = rng.random((5, op['N']))
    avg = onp.avg_populations(y, op['index_S_i'], op['index_A_j'], 2, 7)
    assert avg.shape == (9, 5)
    assert np.allclose(avg[1], (y[:, 8] + y[:, 10]) / 2)
    assert np.allclose(avg[2 + 3], (y[:, 3] + y[:, 14]) / 2)
    t = np.array([0, 1, 3, 6, 10.0])
    cs = onp.cumulative_steady_state(t, y)
    from scipy.integrate import cumulative_trapezoid
    assert np.allclose(cs, cumulative_trapezoid(y, x=t, axis=0, initial=0))


def test_c2_golden_is_self_consistent():
    """The full-size golden trajectories of the bench workload (`oracle/make_golden_c2.py`): ion
    populations are conserved (every operator's columns sum to zero), and the stored emission
    curves are the ion averages of the stored state vectors (`simulations.py:100-128`)."""
    import os
    from tests.helpers import GOLDEN
    with np.load(os.path.join(GOLDEN, 'c2_power_sweep.npz')) as z:
        gold = {k: z[k] for k in z.files}
    ion_type = gold['ion_type']
    iS, iA, pop = onp.ground_states(ion_type, 2, 9)
    assert gold['y_rows'].shape == (3, len(gold['rows']), len(pop))
    for slot in range(3):
        y = gold['y_rows'][slot]
        for gs_list, n in ((iS, 2), (iA, 9)):
            for gs in np.asarray(gs_list).ravel():
                if gs != -1:
                    assert np.allclose(y[:, gs:gs + n].sum(axis=1), 1.0, atol=1e-9)
        avg = onp.avg_populations(y, iS, iA, 2, 9).T
        assert np.allclose(avg, gold['avg'][slot][gold['rows']], rtol=1e-12, atol=1e-300)
    # more power, more excitation: the Yb excited-state population right after the pulse grows
    exc = [gold['y_pulse_end'][s][np.asarray(iS).ravel()[np.asarray(iS).ravel() != -1] + 1].mean() for s in range(3)]
    assert exc[0] < exc[1] < exc[2]


def test_row_sampled_operators_match_the_full_ones():
    """`oracle_np.rate_eq_rows` (the O(n)-per-ion check used for lattices whose reference operands
    do not fit in memory) against the full reference-layout RHS and Jacobian-vector product."""
    from simetuc_b200 import configs, lattice
    cte = configs.yb_tm_cte(N_uc=6, S_conc=12.0, A_conc=2.0, power_dens=1e7)
    lat = lattice.generate_arrays(cte, seed=77)
    L = lat['lattice']
    g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'], 'info': np.array([2, 9])}
    op = oracle_operands(g, cte)
    rng = np.random.default_rng(3)
    y = 10.0 ** rng.uniform(-12, 0, op['N'])
    v = rng.standard_normal(op['N'])
    cM, cN, _ = op['coop']
    f_full = onp.rate_eq_pulse(0.0, y, op['abs'], op['decay'], op['ET'], op['N_indices'], cM, cN)
    jv_full = onp.jac_vec(y, v, (op['abs'] + op['decay']).tocsr(), op['ET'], op['N_indices'])
    aS, aA = onp.absorption_blocks(2, 9, cte['excitations'])
    dS, dA = onp.decay_blocks(2, 9, cte['decay'])
    ions = rng.choice(len(lat['ion_type']), size=12, replace=False)
    idx, f_rows, jv_rows = onp.rate_eq_rows(ions, lat['ion_type'], L.positions, L.cell, 2, 9,
                                            pair_processes(cte), aS + dS, aA + dA, y, v)
    assert np.all(np.abs(f_rows - f_full[idx]) <= 1e-13 * np.abs(f_full).max())
    assert np.all(np.abs(jv_rows - jv_full[idx]) <= 1e-13 * np.abs(jv_full).max())
