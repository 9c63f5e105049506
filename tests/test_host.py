"""CPU-side tests: C-ABI exports, host logic (settings, lattice, blocks), multi-process plumbing."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_np as onp
from simetuc_b200 import configs, distributed, lattice, precalculate, settings, util
from simetuc_b200.lattice import Lattice

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports every function `include/*.h` declares."""
    header = open(os.path.join(ROOT, 'include', 'simetuc_b200.h')).read()
    declared = set(re.findall(r'\b(stb_[a-z_]+)\s*\(', header))
    from simetuc_b200 import _lib
    lib = _lib.load()
    assert declared, 'no declarations parsed'
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_lib.EXPORTS) <= declared
    buf = ctypes.create_string_buffer(512)
    assert lib.stb_describe(buf, 512) == 0 and b'sm_100a' in buf.value
    assert ctypes.sizeof(_lib.StbProcess) == 40 and ctypes.sizeof(_lib.StbStats) == 80


def test_no_cpu_fallback_without_gpu():
    from simetuc_b200 import _lib
    if _lib.load().stb_device_count() > 0:
        pytest.skip('GPU present')
    cte = configs.sim_cte()
    lat = Lattice.from_arrays([1, 0, 0, 1], np.ones((4, 4)) - np.eye(4), 2, 7)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        precalculate.build_system(cte, lat)


def test_blocks_match_oracle():
    for cte in (configs.test_cte(), configs.sim_cte(), configs.yb_tm_cte(), configs.esa_cte(),
                configs.yb_tm_cte(activator_states=7)):
        nS, nA = cte['states']['sensitizer_states'], cte['states']['activator_states']
        aS, aA = precalculate.absorption_blocks(cte)
        oS, oA = onp.absorption_blocks(nS, nA, cte['excitations'])
        assert np.array_equal(aS, oS) and np.array_equal(aA, oA)
        dS, dA = precalculate.decay_blocks(cte)
        pS, pA = onp.decay_blocks(nS, nA, cte['decay'])
        for mine, ref in ((dS, pS), (dA, pA)):
            assert (mine is None and ref is None) or np.array_equal(mine, ref)
        assert sorted(precalculate.get_lifetimes(cte)) == sorted(onp.lifetimes(cte['decay']))
    D = precalculate.decay_blocks(configs.test_cte())[1]
    assert np.allclose(D.sum(axis=0), 0) and np.allclose(D, np.triu(D))


def test_lattice_generator_properties():
    cte = configs.yb_tm_cte(N_uc=8, S_conc=10.0, A_conc=1.0)
    a = lattice.generate(cte, seed=5, with_distances=True)
    b = lattice.generate(cte, seed=5, with_distances=True)
    assert np.array_equal(a.ion_type, b.ion_type) and np.array_equal(a.positions, b.positions)
    assert a.ion_type.dtype == np.uint32
    d = a.dist_array
    assert d.shape == (a.n_ions, a.n_ions) and np.array_equal(d, d.T) and np.all(np.diag(d) == 0)
    assert d[d > 0].min() > 3.4                          # nearest Y-Y distance in beta-NaYF4
    assert d.max() <= 0.5 * np.linalg.norm(a.cell, axis=1).sum() + 1e-9   # minimum image
    iS, iA, pop = a.ground_states()
    oS, oA, opop = onp.ground_states(a.ion_type, 2, 9)
    assert np.array_equal(iS, oS) and np.array_equal(iA, oA) and np.array_equal(pop, opop)
    assert a.n_states == 2 * a.n_S + 9 * a.n_A
    frac = (a.n_S + a.n_A) / (1.5 * 8 ** 3)
    assert 0.07 < frac < 0.16                            # ~11 % of ~768 sites doped
    c = lattice.generate(cte, seed=6)
    assert not np.array_equal(a.ion_type, c.ion_type) or a.n_ions != c.n_ions


def test_lattice_save_load_roundtrip(tmp_path):
    cte = configs.yb_tm_cte(N_uc=5, S_conc=20.0, A_conc=5.0)
    a = lattice.generate(cte, seed=1, with_distances=True)
    path = str(tmp_path / 'lat.npz')
    a.save(path)
    b = Lattice.load(path)
    assert np.array_equal(a.ion_type, b.ion_type) and np.array_equal(a.dist_array, b.dist_array)
    assert (b.n_S_states, b.n_A_states, b.N_uc, b.seed) == (2, 9, 5, 1)


@pytest.mark.parametrize('bad', [dict(S_conc=-1.0), dict(S_conc=60.0, A_conc=50.0), dict(N_uc=0)])
def test_lattice_errors(bad):
    cte = configs.yb_tm_cte()
    cte['lattice'].update(bad)
    with pytest.raises(util.LatticeError):
        lattice.generate(cte, seed=1)


def test_lattice_error_no_ions_and_zero_states():
    cte = configs.yb_tm_cte(N_uc=1, S_conc=0.0, A_conc=0.0)
    with pytest.raises(util.LatticeError):
        lattice.generate(cte, seed=1)
    cte = configs.yb_tm_cte()
    cte['states']['activator_states'] = 0
    with pytest.raises(util.LatticeError):
        lattice.generate(cte, seed=1)


def test_settings_helpers():
    p = settings.parse_power_dependence([1e1, 1e7, 7])
    assert np.allclose(p, [1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7])
    c = settings.parse_conc_dependence([[0], [0.1, 0.3, 0.5, 1.0]], 20, [65, 40, 35, 25])
    assert c['concentrations'] == [(0.0, 0.1), (0.0, 0.3), (0.0, 0.5), (0.0, 1.0)]
    c = settings.parse_conc_dependence([[1, 5], [0.1, 0.3]], 20, [10])
    assert c['concentrations'] == [(1.0, 0.1), (5.0, 0.1), (1.0, 0.3), (5.0, 0.3)]
    assert c['N_uc_list'] == [10, 10, 10, 10]
    cte = settings.Settings.load_from_dict({'lattice': {}})
    assert cte.simulation_params == {'rtol': 1e-5, 'atol': 1e-20, 'N_steps_pulse': 2, 'N_steps': 1000}
    assert cte.lattice == {}


def test_domain_types():
    T, P = util.Transition, util.EneryTransferProcess
    p = P([T(util.IonType.S, 1, 0), T(util.IonType.A, 0, 2)], mult=6, strength=1e4)
    assert p.indices == (1, 0, 0, 2) and [util.ion_code(t) for t in p.type] == [0, 1]
    p.value = 3.0
    assert p.strength == 3.0 and not p.cooperative
    d = util.DecayTransition(util.IonType.A, 3, 1, branching_ratio=0.3)
    d.value = 0.5
    assert d.branching_ratio == 0.5 and d == T(util.IonType.A, 3, 1)
    with pytest.raises(ValueError):
        util.DecayTransition(util.IonType.A, 3, 1)


def test_process_state_check():
    cte = configs.test_cte()
    cte['states']['activator_states'] = 4
    with pytest.raises(util.LatticeError):
        precalculate.check_process_states(cte)


def test_round_robin_sharding():
    for n, w in ((32, 8), (7, 2), (3, 4)):
        shards = [distributed.shard_round_robin(n, w, r) for r in range(w)]
        assert sorted(i for s in shards for i in s) == list(range(n))
        data = np.arange(n, dtype=np.float64)[:, None] * np.ones((1, 3))
        back = distributed.unshard_round_robin([data[s] for s in shards], n)
        assert np.array_equal(back, data)


GLOO_SCRIPT = r'''
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as dist
from simetuc_b200 import distributed
dist.init_process_group('gloo')
rank, size, _ = distributed.world()
ids = distributed.shard_round_robin(7, size, rank)
local = np.array([[i, 10.0 * i] for i in ids], dtype=np.float64).reshape(len(ids), 2)
gathered = distributed.gather_results(local)
tmax = distributed.max_over_ranks([1.0 + rank, 5.0 - rank])
assert tmax == [float(size), 5.0], tmax
if rank == 0:
    full = distributed.unshard_round_robin(gathered, 7)
    assert np.array_equal(full[:, 0], np.arange(7)) and np.array_equal(full[:, 1], 10.0 * np.arange(7))
    print('GATHER_OK')
else:
    assert gathered is None
dist.destroy_process_group()
'''


def test_two_process_gather_gloo(tmp_path):
    """World-size-2 run of the only cross-rank step of the path (final gather + max timing)."""
    script = tmp_path / 'gloo_job.py'
    script.write_text(GLOO_SCRIPT % ROOT)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', '29533', str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'GATHER_OK' in res.stdout


def test_average_lattice_and_processes():
    """Average rate equations (`precalculate.py:884-1051`): a two-ion lattice whose distances,
    self-distances included, are all 1; processes at `strength_avg` without touching the caller's."""
    from simetuc_b200 import precalculate
    from simetuc_b200.util import LatticeError
    cte = configs.yb_tm_cte(N_uc=8, S_conc=10.0, A_conc=1.0)
    lat = precalculate.average_lattice(cte)
    assert (lat.n_ions, lat.n_S, lat.n_A, lat.n_states) == (2, 1, 1, 11)
    assert np.array_equal(lat.dist_array, np.ones((2, 2)))
    cte['lattice']['S_conc'] = 0.0
    lat = precalculate.average_lattice(cte)
    assert (lat.n_ions, lat.n_S, lat.n_A) == (1, 0, 1)
    cte['lattice']['A_conc'] = 0.0
    with pytest.raises(LatticeError):
        precalculate.average_lattice(cte)
    cte['lattice']['A_conc'] = 120.0
    with pytest.raises(LatticeError):
        precalculate.average_lattice(cte)
    cte = configs.yb_tm_cte(N_uc=8, S_conc=10.0, A_conc=1.0)
    first = next(iter(cte['energy_transfer'].values()))
    strength = first.strength
    first._strength_avg = 42.0
    avg = precalculate.average_processes(cte)
    assert next(iter(avg.values())).strength == 42.0
    assert first.strength == strength and list(avg) == list(cte['energy_transfer'])
    # the oracle's interaction lists on the all-ones distance matrix contain the ion itself
    iS, iA, pop = onp.ground_states(np.array([0, 1]), 2, 9)
    lists = onp.interaction_lists(np.array([0, 1]), np.ones((2, 2)), iS, iA, 2, 9)
    assert [list(r) for r in lists[0]] == [[0]] and [list(r) for r in lists[3]] == [[2]]
    assert np.array_equal(precalculate.jac_indices_of(np.array([[1, 5], [3, 7]], np.uint32)),
                          onp.jac_indices(np.array([[1, 5], [3, 7]], np.uint32)))


def test_objective_pieces_match_the_reference(tmp_path, monkeypatch):
    """`DynamicsSolution._load_exp_data / _get_background / list_interp_data / errors` against the
    UNMODIFIED reference's (`simulations.py:354-527`) on the four decay curves shipped for the
    config_file.cfg system; vectors made by `oracle/make_golden_expdata.py`."""
    from simetuc_b200 import simulations
    with np.load(os.path.join(ROOT, 'tests', 'golden', 'expdata_objective.npz')) as z:
        gold = {k: z[k] for k in z.files}
    states = [str(s) for s in gold['states']]
    exp_dir = tmp_path / 'expData' / 'bNaYF4'
    exp_dir.mkdir(parents=True)
    for state in states:
        np.savetxt(exp_dir / ('decay_%s_exc_NIR_980_25.0Yb_0.3Tm.txt' % state), gold['raw_' + state],
                   fmt='%.18e', delimiter='\t', header='t counts')
    monkeypatch.chdir(tmp_path)
    DS = simulations.DynamicsSolution
    for state in states:
        exp = DS._load_exp_data('decay_%s_exc_NIR_980_25.0Yb_0.3Tm.txt' % state, 'bNaYF4')
        assert exp.shape == gold['exp_' + state].shape
        assert np.allclose(exp, gold['exp_' + state], rtol=1e-12, atol=1e-15)
        sim = gold['sim_' + state]
        assert np.isclose(DS._get_background(exp, sim), gold['offset_' + state], rtol=1e-12)
    # the whole chain through a solution object: Yb (GS, ES) + 9 Tm states, curves in state order
    cte = configs.yb_tm_cte(N_uc=12, S_conc=25.0, A_conc=0.3)
    labels = ['Yb_GS', 'Yb_ES'] + ['Tm_' + s for s in cte['states']['activator_states_labels']]
    t_sol = gold['t_sol']
    curves = np.stack([gold['sim_' + lab] if lab in states else np.full(len(t_sol), 1e-3)
                       for lab in labels], axis=1)
    sol = DS(t_sol, curves, [0, -1], [-1, 2], cte)
    sol.set_avg_data(curves)
    assert sol.state_labels == labels
    for k, lab in enumerate(labels):
        if lab in states:
            assert np.allclose(sol.list_interp_data[k], gold['interp_' + lab], rtol=1e-12, atol=1e-300)
            assert np.isclose(sol.errors[k], gold['err_' + lab], rtol=1e-10)
        else:
            assert sol.errors[k] == 0.0
