"""Build-container-only pins of the oracle against the live reference tree.

1. The reference's OWN hand-checked known-answer tests
   (`simetuc/test/test_setup/test_setup.py:25-598`, "THESE RESULTS HAVE BEEN CHECKED BY HAND")
   are executed with `oracle.setup_eqs.setup_microscopic_eqs` substituted for the reference
   builder (`oracle/ref_pytest_plugin.py`): abs / decay / ET matrices, `N_indices`,
   `jac_indices` and the cooperative operands of 1A, 2A, 1S, 2S, 1S1A, 2S2A all match.
2. Oracle builders vs the unmodified reference functions on a seeded synthetic lattice.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.reference


def test_reference_known_answer_tests_pass_with_oracle(tmp_path):
    ini = tmp_path / 'pytest.ini'
    ini.write_text('[pytest]\n')
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=ROOT)
    cmd = [sys.executable, '-m', 'pytest', '-c', str(ini), '--rootdir', str(tmp_path),
           '-p', 'no:cacheprovider', '-p', 'oracle.ref_pytest_plugin',
           '/root/reference/simetuc/test/test_setup/test_setup.py', '-k', 'test_lattice_', '-q']
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT)
    tail = res.stdout[-2000:]
    assert res.returncode == 0, tail
    assert '8 passed' in tail, tail


def test_oracle_matches_live_reference_on_synthetic_lattice():
    from oracle import ref_import, oracle_np as onp
    from simetuc_b200 import configs, lattice
    ref = ref_import.load()
    cte = configs.yb_tm_cte(N_uc=6, S_conc=20.0, A_conc=4.0, types=ref.util)
    lat = lattice.generate_arrays(cte, seed=11)
    sS, sA = 2, 9
    iS, iA, pop = onp.ground_states(lat['ion_type'], sS, sA)
    info = {'num_total': len(lat['ion_type']), 'energy_states': len(pop),
            'activator_states': sA, 'sensitizer_states': sS}
    rS, rA, rpop = ref.lattice.create_ground_states(lat['ion_type'], info)
    assert np.array_equal(rS.ravel(), iS) and np.array_equal(rA.ravel(), iA)
    assert np.array_equal(rpop, pop)
    mine = onp.interaction_lists(lat['ion_type'], lat['dist_array'], iS, iA, sS, sA)
    theirs = ref.lattice.create_interaction_matrices(lat['ion_type'], lat['dist_array'], rS, rA, info)
    for a, b in zip(mine, theirs):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x).ravel(), np.asarray(y).ravel())
    et = {k: v for k, v in cte['energy_transfer'].items() if len(v.type) == 2}
    ET, N_idx = onp.ET_matrices(iS, iA, et, *mine, sS, sA)
    rET, rN = ref.precalculate._create_ET_matrices(
        list(iS), list(iA), et, *[[np.asarray(r).ravel() for r in lst] for lst in theirs[:4]],
        *[np.array([np.asarray(r).ravel() for r in lst]) for lst in theirs[4:]], sS, sA)
    assert np.array_equal(N_idx, rN)
    assert abs(ET - rET).max() == 0
    assert np.array_equal(onp.jac_indices(N_idx), ref.precalculate._calculate_jac_matrices(rN))
    a1 = onp.total_absorption_matrix(sS, sA, cte['excitations'], iS, iA)
    a2 = ref.precalculate._create_total_absorption_matrix(sS, sA, len(pop), cte['excitations'],
                                                          list(iS), list(iA))
    assert abs(a1 - a2).max() == 0
    d1 = onp.decay_matrix(sS, sA, cte['decay'], iS, iA)
    d2 = ref.precalculate._create_decay_matrix(sS, sA, cte['decay'], list(iS), list(iA))
    assert abs(d1 - d2).max() == 0
