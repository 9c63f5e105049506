"""Oracle restatement of `precalculate.setup_microscopic_eqs` (`precalculate.py:713-878`)
for lattices stored in the reference's HDF5 layout (`lattice.py:527-548`) or given as arrays.

TEST INFRASTRUCTURE.  Used to (a) substitute the oracle into the reference's own known-answer
tests (`tests/test_oracle_vs_reference.py`, build container only) and (b) produce operands for
parity tests from `tests/golden/*.npz` lattices.
"""
import re

import numpy as np

from . import oracle_np as onp


def lattice_arrays_from_hdf5(path):
    """Decode a reference lattice file into a dict of arrays (+ parsed `lattice_info`)."""
    from .hdf5_min import load
    data, attrs = load(path)
    info = {k: int(v) for k, v in re.findall(r'(\w+):\s*(\d+)', attrs['lattice_info'])}
    data['lattice_info'] = info
    return data


def _ragged(arr, n_rows):
    """HDF5 `(n, m, 1)` / `(n, m)` / empty datasets -> list of n 1-D arrays."""
    if arr is None or arr.size == 0:
        return [np.zeros(0) for _ in range(n_rows)]
    arr = np.asarray(arr)
    return [arr[r].ravel() for r in range(arr.shape[0])]


def operands_from_lattice(lat, cte, d_max_coop=None):
    """Build the 11 numerical members of the reference's 12-tuple from lattice arrays `lat`
    (keys as in the HDF5 file) and the settings `cte` (`precalculate.py:794-878`)."""
    info = lat['lattice_info']
    sS, sA = info['sensitizer_states'], info['activator_states']
    nS, nA = info['num_sensitizers'], info['num_activators']
    index_S_i = [int(v) for v in np.asarray(lat['indices_S_i']).ravel()]
    index_A_j = [int(v) for v in np.asarray(lat['indices_A_j']).ravel()]
    iSk = [r.astype(np.int64) for r in _ragged(lat.get('index_S_k'), nS)]
    iSl = [r.astype(np.int64) for r in _ragged(lat.get('index_S_l'), nS)]
    iAk = [r.astype(np.int64) for r in _ragged(lat.get('index_A_k'), nA)]
    iAl = [r.astype(np.int64) for r in _ragged(lat.get('index_A_l'), nA)]
    dSk, dSl = _ragged(lat.get('dist_S_k'), nS), _ragged(lat.get('dist_S_l'), nS)
    dAk, dAl = _ragged(lat.get('dist_A_k'), nA), _ragged(lat.get('dist_A_l'), nA)
    initial_population = np.asarray(lat['initial_population'])

    bad = onp.check_process_states(cte['energy_transfer'], sS, sA)
    if bad is not None:
        raise ValueError('The number of A or S states is lower than required by process %s.' % bad)
    abs_m = onp.total_absorption_matrix(sS, sA, cte['excitations'], index_S_i, index_A_j)
    dec_m = onp.decay_matrix(sS, sA, cte['decay'], index_S_i, index_A_j)
    ET, N_idx = onp.ET_matrices(index_S_i, index_A_j, cte['energy_transfer'],
                                iSk, iSl, iAk, iAl, dSk, dSl, dAk, dAl, sS, sA)
    jac = onp.jac_indices(N_idx)
    if d_max_coop is None:
        d_max_coop = cte['lattice'].get('d_max_coop', np.inf)
    cET, cN = onp.coop_ET_matrices(index_S_i, index_A_j, cte['energy_transfer'],
                                   iAk, dAk, dSk, sS, sA, d_max_coop=d_max_coop)
    cJ = onp.coop_jac_indices(cN)
    return (initial_population, index_S_i, index_A_j, abs_m, dec_m, ET, N_idx, jac, cET, cN, cJ)


def setup_microscopic_eqs(cte, gen_lattice=False, full_path=None):
    """Same signature and 12-tuple as `precalculate.setup_microscopic_eqs`
    (`precalculate.py:713-717, 875-878`); `full_path` must name an existing lattice file."""
    lat = lattice_arrays_from_hdf5(full_path)
    info = lat['lattice_info']
    cte['ions'] = {'total': info['num_total'], 'sensitizers': info['num_sensitizers'],
                   'activators': info['num_activators']}
    cte['states']['energy_states'] = info['energy_states']
    cte['states']['sensitizer_states'] = info['sensitizer_states']
    cte['states']['activator_states'] = info['activator_states']
    return (cte,) + operands_from_lattice(lat, cte)
