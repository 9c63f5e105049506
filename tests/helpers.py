"""Shared helpers for the parity tests: golden fixtures -> oracle operands."""
import os

import numpy as np
import scipy.sparse as sp

from oracle import oracle_np as onp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SMALL = ['0S_1A', '0S_2A', '1S_0A', '1S_1A', '2S_0A', '2S_2A', '3S_1A', '3S_2A', 'sim_2S_2A']


def load_golden(tag):
    with np.load(os.path.join(GOLDEN, 'lattice_%s.npz' % tag), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def pair_processes(cte):
    return {k: v for k, v in cte['energy_transfer'].items() if len(v.type) == 2}


def oracle_operands(g, cte, coop=False):
    """Oracle path from `(ion_type, dist_array)` to the reference's operand set."""
    sS, sA = int(g['info'][0]), int(g['info'][1])
    iS, iA, pop = onp.ground_states(g['ion_type'], sS, sA)
    lists = onp.interaction_lists(g['ion_type'], g['dist_array'], iS, iA, sS, sA)
    et = cte['energy_transfer'] if coop else pair_processes(cte)
    abs_m = onp.total_absorption_matrix(sS, sA, cte['excitations'], iS, iA)
    dec_m = onp.decay_matrix(sS, sA, cte['decay'], iS, iA)
    ET, N_idx = onp.ET_matrices(iS, iA, et, *lists, sS, sA)
    jac = onp.jac_indices(N_idx)
    N = len(pop)
    if coop:
        cM, cN = onp.coop_ET_matrices(iS, iA, et, lists[2], lists[6], lists[4], sS, sA,
                                      d_max_coop=cte['lattice'].get('d_max_coop', np.inf))
        cJ = onp.coop_jac_indices(cN)
    else:
        cM, cN, cJ = onp.empty_coop(N)
    return dict(sS=sS, sA=sA, index_S_i=iS, index_A_j=iA, initial_population=pop, lists=lists,
                abs=abs_m, decay=dec_m, ET=ET, N_indices=N_idx, jac_indices=jac,
                coop=(cM, cN, cJ), N=N)
