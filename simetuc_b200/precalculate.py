"""Equation setup for the microscopic rate equations - host mirror of
`simetuc/precalculate.py`.

The per-ion-type absorption and decay blocks (`precalculate.py:79-230`) are tiny (at most
16x16) and are computed here on the host; the O(n_ions^2) part - the interaction lists of
`lattice.create_interaction_matrices` and the pair rates of `_create_ET_matrices`
(`precalculate.py:235-415`) - is built on the device by `RateSystem`.
`setup_microscopic_eqs` keeps the reference's signature and 12-tuple; the sparse matrices it
returns are materialised from the device export for API compatibility and carry a reference to
the device-resident `RateSystem`, which `odesolver.solve_pulse / solve_relax` use directly.
"""
import collections
import copy
import logging
import os
import threading
import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp

from . import lattice as lattice_mod
from .engine import RateSystem
from .lattice import Lattice
from .util import LatticeError, ion_code

# in-memory lattice cache: the reference caches lattices on disk under latticeData/
# (`lattice.py:377-385`, `precalculate.py:755-791`)
_LATTICE_CACHE: Dict[Tuple, Lattice] = {}
_SYSTEM_CACHE: Dict[Tuple, RateSystem] = {}
# device-resident systems of finished sweeps, parked for the next sweep on the same lattice with the
# same process table (a sweep CHECKS a system OUT: no two threads ever share a handle).  The
# reference rebuilds every matrix for every power / candidate (`simulations.py:1118,1214`).
_SYSTEM_POOL: 'collections.OrderedDict[Tuple, List[RateSystem]]' = collections.OrderedDict()
_SYSTEM_POOL_LOCK = threading.Lock()
SYSTEM_POOL_MAX = 8                      # parked handles (least recently used are closed)
SYSTEM_POOL_MAX_BYTES = 24 * 2 ** 30     # of pair data, all parked handles together


def absorption_blocks(cte) -> Tuple[np.ndarray, np.ndarray]:
    """Per-type absorption blocks (already multiplied by the power density), summed over the
    processes of every active excitation (`precalculate.py:79-127`)."""
    nS, nA = cte['states']['sensitizer_states'], cte['states']['activator_states']
    blocks = [np.zeros((nS, nS)), np.zeros((nA, nA))]
    for exc_list in cte['excitations'].values():
        for exc in exc_list:
            if exc.active is False:
                continue
            t = ion_code(exc.transition.ion)
            n = (nS, nA)[t]
            i, f = exc.transition.state_i, exc.transition.state_f
            if n == 0 or i >= n or f >= n:
                continue
            one = np.zeros((n, n))
            one[i, i] = -exc.pump_rate
            one[i, f] = exc.degeneracy * exc.pump_rate
            one[f, i] = exc.pump_rate
            one[f, f] = -exc.degeneracy * exc.pump_rate
            blocks[t] = blocks[t] + one * exc.power_dens
    return blocks[0], blocks[1]


def decay_blocks(cte) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """Per-type decay blocks `(B^T - I) diag(k)` with row 0 = -column sums
    (`precalculate.py:130-213`; branching to the ground state dropped `:153-156`)."""
    nS, nA = cte['states']['sensitizer_states'], cte['states']['activator_states']
    decay = cte['decay']
    out = []
    for t, n, bkey, kkey in ((0, nS, 'branching_S', 'decay_S'), (1, nA, 'branching_A', 'decay_A')):
        if n == 0:
            out.append(None)
            continue
        B = np.zeros((n, n))
        for tr in decay[bkey]:
            # the reference's activator check tests state_i twice (`precalculate.py:146`)
            inside = (tr.state_i < n and tr.state_f < n) if t == 0 else tr.state_i < n
            if inside:
                B[tr.state_i, tr.state_f] = tr.branching_ratio
        B[:, 0] = 0.0
        k = np.zeros(n)
        for tr in decay[kkey]:
            if tr.state_i < n:
                k[tr.state_i] = tr.decay_rate
        D = (B.T - np.eye(n)) * k[None, :]
        D[0, :] = -D.sum(axis=0)
        out.append(D)
    return out[0], out[1]


def get_lifetimes(cte) -> List[float]:
    """All lifetimes in seconds (`precalculate.py:702-709`)."""
    return [1 / float(tr.decay_rate) for tr in list(cte['decay']['decay_S']) + list(cte['decay']['decay_A'])]


def _block_diag(block_S, block_A, index_S_i, index_A_j):
    blocks = []
    for s, a in zip(index_S_i, index_A_j):
        if s != -1 and block_S is not None and block_S.size:
            blocks.append(sp.csr_matrix(block_S))
        elif s == -1 and a != -1 and block_A is not None and block_A.size:
            blocks.append(sp.csr_matrix(block_A))
    if not blocks:
        return sp.csr_matrix((0, 0), dtype=np.float64)
    return sp.block_diag(blocks, format='csr', dtype=np.float64)


def _lattice_key(cte):
    lat = cte['lattice']
    return (lat.get('name'), int(lat['N_uc']), float(lat['S_conc']), float(lat['A_conc']),
            lat.get('radius'), cte['states']['sensitizer_states'],
            cte['states']['activator_states'], cte.get('lattice_seed'))


def get_lattice(cte, gen_lattice: bool = False, full_path: Optional[str] = None) -> Lattice:
    """Lattice for `cte`: a file written by `Lattice.save` (`full_path`), the in-memory cache,
    or a fresh realisation (`gen_lattice=True` or cache miss) seeded by `cte['lattice_seed']`."""
    if full_path is not None:
        if not os.path.exists(full_path):
            raise FileNotFoundError(full_path)
        lat = Lattice.load(full_path)
        cte['states']['sensitizer_states'] = lat.n_S_states
        cte['states']['activator_states'] = lat.n_A_states
        return lat
    key = _lattice_key(cte)
    if gen_lattice or key not in _LATTICE_CACHE:
        _LATTICE_CACHE[key] = lattice_mod.generate(cte, seed=cte.get('lattice_seed'))
    return _LATTICE_CACHE[key]


def check_process_states(cte):
    """`precalculate.py:325-345`: an active pair process naming more states than exist raises
    `LatticeError` (the reference compares with `>`; so does this)."""
    nS, nA = cte['states']['sensitizer_states'], cte['states']['activator_states']
    count = {0: nS, 1: nA}
    for name, p in cte['energy_transfer'].items():
        if len(p.type) != 2 or np.isclose(p.strength, 0.0):
            continue
        ii, fi, if_, ff = p.indices
        d, a = ion_code(p.type[0]), ion_code(p.type[1])
        if max(ii, if_) > count[d] or max(fi, ff) > count[a]:
            raise LatticeError('The number of A or S states is lower than required by '
                               'process {}.'.format(name))


def _system_key(lat: Lattice, processes, d_max_coop, device: int, d_max=None) -> Tuple:
    """What the device-resident pair data depends on: the lattice object, which processes exist and
    are active (strengths of the pair processes are per-system parameters; the cooperative one is
    baked into the handle), the cutoffs and the device."""
    return (id(lat), tuple((n, tuple(p.indices), p.mult, bool(np.isclose(p.strength, 0.0)),
                            p.strength if len(p.type) == 3 else None)
                           for n, p in processes.items()), d_max_coop, d_max, device)


def acquire_system(cte, lat: Lattice, device: int = 0, n_sys: int = 1, processes=None,
                   d_max: Optional[float] = None) -> RateSystem:
    """Check a device-resident system for `lat` out of the pool (or build one): the pair build and
    its device allocations are paid once per lattice, not once per sweep.  Hand it back with
    `release_system`.  The batch is resized to `n_sys`; blocks and strengths are the caller's to
    set."""
    processes = cte['energy_transfer'] if processes is None else processes
    d_max_coop = cte['lattice'].get('d_max_coop') if hasattr(cte['lattice'], 'get') else None
    key = _system_key(lat, processes, d_max_coop, device, d_max)
    system = None
    with _SYSTEM_POOL_LOCK:
        parked = _SYSTEM_POOL.get(key)
        if parked:
            system = parked.pop()
            if not parked:
                del _SYSTEM_POOL[key]
    if system is None or system.handle is None:
        check_process_states(cte)
        system = RateSystem(lat, processes, device=device, d_max=d_max, d_max_coop=d_max_coop)
        system.pool_key = key
    system.active_params = None
    if system.n_sys != n_sys:
        system.set_batch(n_sys)
    return system


def release_system(system: RateSystem) -> None:
    """Park a checked-out system for the next sweep on its lattice; the least recently used
    handles are closed beyond `SYSTEM_POOL_MAX` / `SYSTEM_POOL_MAX_BYTES`."""
    key = getattr(system, 'pool_key', None)
    if key is None or system.handle is None:
        system.close()
        return
    evict = []
    with _SYSTEM_POOL_LOCK:
        _SYSTEM_POOL.setdefault(key, []).append(system)
        _SYSTEM_POOL.move_to_end(key)

        def total():
            return sum(len(v) for v in _SYSTEM_POOL.values()), \
                sum(s.stream_bytes for v in _SYSTEM_POOL.values() for s in v)
        count, nbytes = total()
        while _SYSTEM_POOL and (count > SYSTEM_POOL_MAX or nbytes > SYSTEM_POOL_MAX_BYTES):
            old_key = next(iter(_SYSTEM_POOL))
            lst = _SYSTEM_POOL[old_key]
            evict.append(lst.pop(0))
            if not lst:
                del _SYSTEM_POOL[old_key]
            count, nbytes = total()
    for old in evict:
        old.close()


def clear_system_pool() -> None:
    """Close every parked device system (frees their device memory)."""
    with _SYSTEM_POOL_LOCK:
        parked = [s for v in _SYSTEM_POOL.values() for s in v]
        _SYSTEM_POOL.clear()
    for system in parked:
        system.close()


class ParamSet:
    """Parameters of ONE operand set returned by `setup_*_eqs`: the per-type blocks and the pair
    strengths its matrices were built from.  Several operand sets may share a device handle
    (`_SYSTEM_CACHE`); `odesolver` re-applies the set of the operands it is given before it solves,
    so operands keep meaning what they meant when they were created."""

    def __init__(self, decay, absorption, strengths):
        self.decay, self.absorption, self.strengths = decay, absorption, strengths

    def apply(self, system: RateSystem, index: int = 0) -> None:
        if getattr(system, 'active_params', None) is self:
            return
        system.set_decay_blocks(*self.decay, system=index)
        system.set_abs_blocks(*self.absorption, system=index)
        system.set_strengths(self.strengths, system=index)
        system.active_params = self


def build_system(cte, lat: Lattice, device: int = 0, d_max: Optional[float] = None,
                 n_sys: int = 1) -> RateSystem:
    """Device-resident system for `lat` with `cte`'s processes; every system of the batch gets
    `cte`'s decay and absorption blocks."""
    check_process_states(cte)
    system = RateSystem(lat, cte['energy_transfer'], device=device, d_max=d_max,
                        d_max_coop=cte['lattice'].get('d_max_coop'))
    system.set_batch(n_sys)
    system.set_decay_blocks(*decay_blocks(cte))
    system.set_abs_blocks(*absorption_blocks(cte))
    return system


def jac_indices_of(N_indices) -> np.ndarray:
    """`[2M,3]` rows `(r, N_i[r], N_j[r])`, `(r, N_j[r], N_i[r])` interleaved
    (`precalculate.py:603-623`)."""
    N_indices = np.asarray(N_indices, dtype=np.uint32).reshape(-1, 2)
    M = N_indices.shape[0]
    out = np.empty((2 * M, 3), dtype=np.uint32)
    out[:, 0] = np.repeat(np.arange(M, dtype=np.uint32), 2)
    out[0::2, 1], out[0::2, 2] = N_indices[:, 0], N_indices[:, 1]
    out[1::2, 1], out[1::2, 2] = N_indices[:, 1], N_indices[:, 0]
    return out


def average_lattice(cte) -> Lattice:
    """Lattice of the AVERAGE rate equations (`precalculate.py:911-977`): one sensitizer and/or one
    activator ion with every distance, the ion's distance to itself included, equal to 1 A - so an
    ion also "interacts with itself", which is how the average equations express A-A and S-S
    processes.  Same checks and errors as the reference (`:930-958`)."""
    lat = cte['lattice']
    S_conc, A_conc = float(lat['S_conc']), float(lat['A_conc'])
    n_S, n_A = int(S_conc != 0), int(A_conc != 0)
    if n_S + n_A == 0:
        raise LatticeError('No ions generated, the concentrations are too small!')
    if lat['N_uc'] <= 0:
        raise LatticeError('Wrong number of unit cells: {}. It must be a positive integer.'.format(lat['N_uc']))
    if not ((0.0 <= S_conc <= 100.0) and (0.0 <= A_conc <= 100.0) and (0 <= S_conc + A_conc <= 100.0)):
        raise LatticeError('Wrong ion concentrations: {:.2f}% Sensitizer, {:.2f}% Activator. '
                           'Their sum must be between 0% and 100%.'.format(S_conc, A_conc))
    sS, sA = cte['states']['sensitizer_states'], cte['states']['activator_states']
    if (S_conc != 0 and sS == 0) or (A_conc != 0 and sA == 0):
        raise LatticeError('The number of states of each ion cannot be zero.')
    ion_type = [0] * n_S + [1] * n_A
    n = len(ion_type)
    return Lattice.from_arrays(ion_type, np.ones((n, n)), sS, sA)


def average_processes(cte):
    """Energy-transfer processes with `strength = strength_avg` where one is given
    (`precalculate.py:999-1003`); the caller's processes are left untouched (the reference
    overwrites them in place through a shallow dict copy)."""
    out = type(cte['energy_transfer'])()
    for name, p in cte['energy_transfer'].items():
        q = copy.copy(p)
        q.strength = p.strength_avg
        out[name] = q
    return out


def average_view(cte):
    """Shallow copy of `cte` whose pair processes carry their average strengths."""
    view = copy.copy(cte)
    view['energy_transfer'] = average_processes(cte)
    return view


def setup_average_eqs(cte, gen_lattice: bool = False, full_path: Optional[str] = None,
                      device: int = 0):
    """Average rate equations (`precalculate.py:884-1051`): the same 12-tuple as
    `setup_microscopic_eqs` for the two-ion average lattice; interaction columns that cancel
    identically (energy migration of an ion with itself) are removed from the exported operands
    as the reference does (`:1013-1021`)."""
    lat = average_lattice(cte)
    out = _setup_eqs(cte, lat, device, average_processes(cte), drop_empty=True)
    logging.getLogger(__name__).info('Number of interactions: %d.', out[7].shape[0])
    return out


def setup_microscopic_eqs(cte, gen_lattice: bool = False, full_path: Optional[str] = None,
                          device: int = 0):
    """Same contract as the reference (`precalculate.py:713-878`): returns
    `(cte, initial_population, index_S_i, index_A_j, total_abs_matrix, decay_matrix, ET_matrix,
    N_indices, jac_indices, coop_ET_matrix, coop_N_indices, coop_jac_indices)`.
    The matrices carry `.stb = (RateSystem, system_index)`."""
    lat = get_lattice(cte, gen_lattice=gen_lattice, full_path=full_path)
    return _setup_eqs(cte, lat, device, cte['energy_transfer'], drop_empty=False)


def _setup_eqs(cte, lat, device, processes, drop_empty):
    logger = logging.getLogger(__name__)
    start = time.time()
    cte['ions'] = {'total': lat.n_ions, 'sensitizers': lat.n_S, 'activators': lat.n_A}
    cte['states']['energy_states'] = lat.n_states
    d_max_coop = (cte['lattice'].get('d_max_coop') if hasattr(cte['lattice'], 'get') else None)
    key = _system_key(lat, processes, d_max_coop, device)
    system = _SYSTEM_CACHE.get(key)
    if system is None or system.handle is None:
        check_process_states(cte)
        system = RateSystem(lat, processes, device=device, d_max_coop=d_max_coop)
        for old_key in [k for k in _SYSTEM_CACHE if k[-1] == device]:
            del _SYSTEM_CACHE[old_key]          # one resident lattice per device is enough here
        _SYSTEM_CACHE[key] = system
    check_process_states(cte)
    system.set_batch(1)
    dS, dA = decay_blocks(cte)
    aS, aA = absorption_blocks(cte)
    params = ParamSet((dS, dA), (aS, aA),
                      np.array([p.strength for n, p in processes.items() if len(p.type) == 2],
                               dtype=np.float64))
    system.active_params = None
    params.apply(system)
    index_S_i, index_A_j, initial_population = system.ground_states()
    total_abs_matrix = _block_diag(aS, aA, index_S_i, index_A_j)
    decay_matrix = _block_diag(dS, dA, index_S_i, index_A_j)
    ET_matrix, N_indices, jac_indices = system.ET_matrix(0)
    if drop_empty and N_indices.shape[0]:
        dense = ET_matrix.toarray()
        keep = np.array([not np.allclose(dense[:, k], 0) for k in range(dense.shape[1])])
        if not keep.all():
            ET_matrix = sp.csr_matrix(dense[:, keep])
            N_indices = N_indices[keep]
            jac_indices = jac_indices_of(N_indices)
    N = system.N
    coop_ET_matrix, coop_N_indices, coop_jac_indices = system.coop_operands()
    for m in (total_abs_matrix, decay_matrix, ET_matrix, coop_ET_matrix):
        m.stb = (system, 0)
        m.stb_params = params
    logger.info('Setup finished. Total time: %.2fs.', time.time() - start)
    return (cte, initial_population, list(index_S_i), list(index_A_j), total_abs_matrix,
            decay_matrix, ET_matrix, N_indices, jac_indices, coop_ET_matrix, coop_N_indices,
            coop_jac_indices)
