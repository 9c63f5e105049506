"""Generate `tests/golden/*.npz` from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

For every lattice fixture the reference ships (`simetuc/test/test_setup/data_*.hdf5`,
`test_simulations/data_2S_2A.hdf5`, `test_benchmark/data_240S_108A.hdf5`) this script stores
  * the lattice arrays exactly as decoded from the HDF5 file (`lattice.py:527-548`), and
  * what the reference's own functions produce from them with the parameter sets of
    `conftest.py:26-117` ("test") and `conftest.py:288-312` ("sim"):
    `_create_total_absorption_matrix`, `_create_decay_matrix`, `_create_ET_matrices`,
    `_calculate_jac_matrices` (`precalculate.py:79-415,603-623`), and `_rate_eq_pulse`,
    `_rate_eq`, `_jac_rate_eq_pulse(y) @ v` (`odesolver.py:33-80`) at seeded random `y`, `v`.
Small fixtures keep full matrices; the 1236-state one keeps vectors plus SHA-256 digests of the
integer lists.  The golden trajectory `t_sol/y_sol_2S_2A.hdf5` is copied verbatim.
The committed files are what travels to the GPU box; `/root/reference` does not.
"""
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import                                   # noqa: E402
from oracle.setup_eqs import lattice_arrays_from_hdf5, _ragged  # noqa: E402

TEST = '/root/reference/simetuc/test'
FIXTURES = {
    '0S_1A': 'test_setup/data_0S_1A.hdf5', '0S_2A': 'test_setup/data_0S_2A.hdf5',
    '1S_0A': 'test_setup/data_1S_0A.hdf5', '1S_1A': 'test_setup/data_1S_1A.hdf5',
    '2S_0A': 'test_setup/data_2S_0A.hdf5', '2S_2A': 'test_setup/data_2S_2A.hdf5',
    '3S_1A': 'test_setup/data_3S_1A.hdf5', '3S_2A': 'test_setup/data_3S_2A.hdf5',
    'sim_2S_2A': 'test_simulations/data_2S_2A.hdf5',
    '240S_108A': 'test_benchmark/data_240S_108A.hdf5',
}


def random_state(rng, index_S_i, index_A_j, sS, sA, N):
    """Log-uniform populations 1e-20..1, normalised per ion (SURVEY.md 8d parity protocol)."""
    y = 10.0 ** rng.uniform(-20, 0, size=N)
    for gs_list, n in ((index_S_i, sS), (index_A_j, sA)):
        for gs in gs_list:
            if gs != -1 and n:
                y[gs:gs + n] /= y[gs:gs + n].sum()
    return y


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = ref_import.load()
    from simetuc_b200 import configs
    pre, ode = ref.precalculate, ref.odesolver
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)

    for tag, rel in FIXTURES.items():
        lat = lattice_arrays_from_hdf5(os.path.join(TEST, rel))
        info = lat['lattice_info']
        sS, sA = info['sensitizer_states'], info['activator_states']
        nS, nA, N = info['num_sensitizers'], info['num_activators'], info['energy_states']
        idx_S = [int(v) for v in lat['indices_S_i'].ravel()]
        idx_A = [int(v) for v in lat['indices_A_j'].ravel()]
        save = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'],
                'doped_lattice': lat['doped_lattice'],
                'indices_S_i': lat['indices_S_i'].ravel(), 'indices_A_j': lat['indices_A_j'].ravel(),
                'initial_population': lat['initial_population'],
                'info': np.array([sS, sA, nS, nA, N], dtype=np.int64)}
        lists = {}
        for key, n in (('S_k', nS), ('S_l', nS), ('A_k', nA), ('A_l', nA)):
            lists['i' + key] = [r.astype(np.int64) for r in _ragged(lat.get('index_' + key), n)]
            lists['d' + key] = _ragged(lat.get('dist_' + key), n)
            # the reference passes dist_* as 2-D arrays (precalculate.py:824-836)
            lists['D' + key] = (np.array(lists['d' + key]) if n else np.zeros((0, 0)))
        big = N > 200
        if not big:
            for key in ('S_k', 'S_l', 'A_k', 'A_l'):
                rows = len(lists['i' + key])
                width = len(lists['i' + key][0]) if rows else 0
                save['index_' + key] = np.array(lists['i' + key], dtype=np.int64).reshape(rows, width)
                save['dist_' + key] = np.asarray(lists['D' + key], dtype=np.float64).reshape(rows, width)

        for pname in ('test', 'sim'):
            cte = getattr(configs, pname + '_cte')(types=ref.util)
            # pair processes only: the coop builder needs numba closures + np.Inf (numpy<2)
            et = {k: v for k, v in cte['energy_transfer'].items() if len(v.type) == 2}
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                abs_m = pre._create_total_absorption_matrix(sS, sA, N, cte['excitations'], idx_S, idx_A)
                dec_m = pre._create_decay_matrix(sS, sA, cte['decay'], idx_S, idx_A)
                ET, N_idx = pre._create_ET_matrices(
                    idx_S, idx_A, et, lists['iS_k'], lists['iS_l'], lists['iA_k'], lists['iA_l'],
                    lists['DS_k'], lists['DS_l'], lists['DA_k'], lists['DA_l'], sS, sA)
            N_idx = N_idx.astype(np.uint32).reshape(-1, 2)
            jac = pre._calculate_jac_matrices(N_idx) if N_idx.shape[0] else np.zeros((0, 3), np.uint32)
            import scipy.sparse as sp
            cM = sp.csr_matrix(np.zeros((N, 0)))
            cN, cJ = np.zeros((0, 3), np.uint32), np.zeros((0, 4), np.uint32)
            rng = np.random.default_rng(20260000 + len(tag) * 7 + (pname == 'sim'))
            ys = np.array([random_state(rng, idx_S, idx_A, sS, sA, N) for _ in range(3)])
            vs = rng.standard_normal((3, N)) * ys
            f_pulse = np.array([ode._rate_eq_pulse(0.0, y, abs_m, dec_m, ET, N_idx, cM, cN) for y in ys])
            f_relax = np.array([ode._rate_eq(0.0, y, dec_m, ET, N_idx, cM, cN) for y in ys])
            if N <= 2000:
                jv = np.array([ode._jac_rate_eq_pulse(0.0, y, abs_m, dec_m, ET, jac, cM, cJ) @ v
                               for y, v in zip(ys, vs)])
            pre_ = pname + '_'
            save.update({pre_ + 'y': ys, pre_ + 'v': vs, pre_ + 'f_pulse': f_pulse,
                         pre_ + 'f_relax': f_relax, pre_ + 'jv_pulse': jv})
            ETc = ET.tocsr()
            ETc.sort_indices()
            if big:
                save.update({pre_ + 'M': np.array([N_idx.shape[0]]),
                             pre_ + 'sha_N_indices': np.array(sha(N_idx)),
                             pre_ + 'sha_jac_indices': np.array(sha(jac.astype(np.uint32))),
                             pre_ + 'sha_ET_indices': np.array(sha(ETc.indices.astype(np.int32))),
                             pre_ + 'sha_ET_indptr': np.array(sha(ETc.indptr.astype(np.int32))),
                             pre_ + 'ET_data_sum_abs': np.array([np.abs(ETc.data).sum()]),
                             pre_ + 'ET_colnorm_head': np.abs(ET.tocsc().data[:4000:4])})
            else:
                save.update({pre_ + 'abs': abs_m.toarray(), pre_ + 'decay': dec_m.toarray(),
                             pre_ + 'ET': ET.toarray(), pre_ + 'N_indices': N_idx,
                             pre_ + 'jac_indices': jac.astype(np.uint32)})
        np.savez_compressed(os.path.join(out_dir, 'lattice_%s.npz' % tag), **save)
        print(tag, 'N=%d' % N, 'M=%d' % N_idx.shape[0])

    # golden trajectory (test_simulations.py:76-91)
    from oracle.hdf5_min import load
    t, _ = load(os.path.join(TEST, 'test_simulations/t_sol_2S_2A.hdf5'))
    y, _ = load(os.path.join(TEST, 'test_simulations/y_sol_2S_2A.hdf5'))
    np.savez_compressed(os.path.join(out_dir, 'sol_2S_2A.npz'), t_sol=t['t_sol'], y_sol=y['y_sol'])
    print('trajectory', y['y_sol'].shape)


if __name__ == '__main__':
    main()
