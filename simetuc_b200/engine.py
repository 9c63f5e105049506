"""`RateSystem`: one doped lattice resident on one B200 plus a batch of parameter sets.

Thin Python owner of the opaque `stb_system` handle of the C ABI (`include/simetuc_b200.h`).
It is what `precalculate.setup_microscopic_eqs` builds and what `odesolver.solve_pulse /
solve_relax` drive; the reference has no such object because it rebuilds every matrix for
every call (`simulations.py:1118,1214`).
"""
import ctypes as C
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp

from . import _lib
from .lattice import Lattice
from .util import LatticeError, ion_code


def pair_processes(energy_transfer) -> list:
    """Two-ion processes of `cte.energy_transfer` in dict order (`precalculate.py:326-391`)."""
    return [(name, p) for name, p in energy_transfer.items() if len(p.type) == 2]


def coop_process(energy_transfer):
    """The first cooperative process of type (S, S, A) with non-zero strength
    (`precalculate.py:512-520`), or None."""
    for p in energy_transfer.values():
        if len(p.type) == 3 and not np.isclose(p.strength, 0.0) and \
                tuple(ion_code(t) for t in p.type) == (0, 0, 1):
            return p
    return None


class RateSystem:
    """Device-resident ET coupling of one lattice + per-system linear blocks and ET strengths."""

    def __init__(self, lattice: Lattice, energy_transfer, device: int = 0,
                 d_max: Optional[float] = None, geometry: str = 'auto',
                 d_max_coop: Optional[float] = None):
        lib = _lib.require_gpu()
        self.lib = lib
        self.lattice = lattice
        self.device = device
        self.proc_names = [name for name, _ in pair_processes(energy_transfer)]
        procs = [p for _, p in pair_processes(energy_transfer)]
        if len(procs) > _lib.STB_MAX_PROC:
            raise ValueError('too many energy transfer processes')
        arr = (_lib.StbProcess * max(len(procs), 1))()
        for k, p in enumerate(procs):
            ii, fi, if_, ff = p.indices
            arr[k] = _lib.StbProcess(ion_code(p.type[0]), ion_code(p.type[1]), ii, fi, if_, ff,
                                     int(p.mult), 0, float(p.strength))
        self._procs = arr
        self.n_proc = len(procs)
        ion_type = np.ascontiguousarray(lattice.ion_type, dtype=np.uint8)
        use_dist = lattice.dist_array is not None if geometry == 'auto' else geometry == 'dist'
        dist = pos = cell = None
        if use_dist:
            dist = _lib.f64(lattice.distances())
        else:
            pos = _lib.f64(lattice.positions)
            cell = None if lattice.cell is None else _lib.f64(lattice.cell)
        self.d_max = -1.0 if d_max is None else float(d_max)
        self.handle = lib.stb_system_create(
            device, len(ion_type), _lib.ptr(ion_type), lattice.n_S_states, lattice.n_A_states,
            _lib.ptr(dist), _lib.ptr(pos), _lib.ptr(cell), self.d_max, self.n_proc,
            C.cast(arr, C.c_void_p))
        if not self.handle:
            msg = _lib.last_error()
            if 'states' in msg or 'process' in msg:
                raise LatticeError(msg)
            raise _lib.StbError(-1, msg)
        info = (C.c_int64 * 8)()
        _lib.check(lib.stb_system_info(self.handle, info))
        (self.n_ions, self.n_S, self.n_A, self.N, self.n_pairs, self.stream_bytes,
         self.n_slots, self.lanes_per_row) = [int(v) for v in info]
        self.sS, self.sA = lattice.n_S_states, lattice.n_A_states
        self.n_sys = 1
        #: bytes streamed per stored pair: 8 (dense all-to-all blocks) or 4 + 8 per multipolarity
        self.bytes_per_pair = self.stream_bytes // self.n_slots if self.n_slots else 0
        self._geometry = (dist, pos, cell)
        self.n_coop = 0
        #: `precalculate.ParamSet` whose blocks / strengths system 0 currently holds (None: unknown)
        self.active_params = None
        #: key under which `precalculate.release_system` parks this handle
        self.pool_key = None
        coop = coop_process(energy_transfer)
        if coop is not None:
            self.set_coop(coop, d_max_coop)

    def set_coop(self, process, d_max_coop: Optional[float] = None):
        """Install (or, with strength 0, remove) the cooperative S-S-A process
        (`precalculate.py:427-599`); call before `set_batch`."""
        dist, pos, cell = self._geometry
        idx6 = np.ascontiguousarray(process.indices, dtype=np.int32)
        if idx6.shape != (6,):
            raise ValueError('a cooperative process has three transitions')
        d_max = -1.0 if d_max_coop is None or not np.isfinite(d_max_coop) else float(d_max_coop)
        rc = self.lib.stb_set_coop(self.handle, _lib.ptr(dist), _lib.ptr(pos), _lib.ptr(cell),
                                   _lib.ptr(idx6), int(process.mult), float(process.strength), d_max)
        if rc != 0:
            msg = _lib.last_error()
            if 'states' in msg or 'process' in msg:
                raise LatticeError(msg)
            raise _lib.StbError(rc, msg)
        self.n_coop = int(self.lib.stb_coop_count(self.handle))

    def coop_operands(self):
        """`(coop_ET_matrix [N, Mc] CSR, coop_N_indices [Mc, 3], coop_jac_indices [3 Mc, 4])` in the
        reference's layout (`precalculate.py:575-599, 626-699`)."""
        Mc = self.n_coop
        if Mc == 0:
            return (sp.csr_matrix((self.N, 0), dtype=np.float64), np.zeros((0, 3), dtype=np.uint32),
                    np.zeros((0, 4), dtype=np.uint32))
        N3 = np.empty((Mc, 3), dtype=np.uint32)
        rows = np.empty((Mc, 6), dtype=np.int32)
        w = np.empty(Mc, dtype=np.float64)
        _lib.check(self.lib.stb_export_coop(self.handle, _lib.ptr(N3), _lib.ptr(rows), _lib.ptr(w)))
        vals = np.column_stack((-w, w, -w, w, -w, w)).ravel()
        cols = np.repeat(np.arange(Mc, dtype=np.int64), 6)
        M = sp.csr_matrix((vals, (rows.ravel().astype(np.int64), cols)), shape=(self.N, Mc), dtype=np.float64)
        jac = np.empty((3 * Mc, 4), dtype=np.uint32)
        jac[:, 0] = np.repeat(np.arange(Mc, dtype=np.uint32), 3)
        jac[0::3, 1], jac[1::3, 1], jac[2::3, 1] = N3[:, 2], N3[:, 1], N3[:, 0]
        jac[0::3, 2], jac[1::3, 2], jac[2::3, 2] = N3[:, 0], N3[:, 0], N3[:, 1]
        jac[0::3, 3], jac[1::3, 3], jac[2::3, 3] = N3[:, 1], N3[:, 2], N3[:, 2]
        return M, N3, jac

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, 'handle', None):
            self.lib.stb_system_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # pragma: no cover - interpreter shutdown
            pass

    def set_option(self, name: str, value: float):
        """`stb_set_option`, e.g. `('sm_budget', 18)`: SMs one solve of this handle may occupy."""
        _lib.check(self.lib.stb_set_option(self.handle, name.encode(), float(value)))

    def set_experiment(self, kind: int, exp_data):
        """Experimental decay curve of state kind `kind` for `objective()`: `exp_data [n_e, 2]` as
        `DynamicsSolution._load_exp_data` returns it (time, normalised counts), or None."""
        if exp_data is None or not np.any(exp_data):
            _lib.check(self.lib.stb_set_experiment(self.handle, int(kind), 0, None, None, 0.0))
            return
        t = _lib.f64(exp_data[:, 0])
        y = _lib.f64(exp_data[:, 1])
        tail = y[y > 0][-50:]                     # `simulations.py:437-441`
        tail_mean = float(np.mean(tail)) if tail.size else 0.0
        _lib.check(self.lib.stb_set_experiment(self.handle, int(kind), len(t), _lib.ptr(t), _lib.ptr(y),
                                               tail_mean))

    def objective(self, n: int):
        """`[n, sS+sA]` deviations of the emission curves of the last `integrate_dynamics` call
        from the experimental curves (`stb_objective`)."""
        out = np.empty((n, self.sS + self.sA))
        _lib.check(self.lib.stb_objective(self.handle, int(n), _lib.ptr(out)))
        return out

    def get_option(self, name: str) -> float:
        """`stb_get_option`: 'sm_budget', 'quad', 'pc_group', 'cell_lists'."""
        out = C.c_double()
        _lib.check(self.lib.stb_get_option(self.handle, name.encode(), C.byref(out)))
        return float(out.value)

    # ------------------------------------------------------------------ lattice-side outputs
    def ground_states(self):
        """`(index_S_i, index_A_j, initial_population)` as `lattice.create_ground_states`."""
        iS = np.empty(self.n_ions, dtype=np.int64)
        iA = np.empty(self.n_ions, dtype=np.int64)
        pop = np.empty(self.N, dtype=np.uint64)
        _lib.check(self.lib.stb_ground_states(self.handle, _lib.ptr(iS), _lib.ptr(iA), _lib.ptr(pop)))
        return iS, iA, pop

    # ------------------------------------------------------------------ parameter batch
    def set_batch(self, n_sys: int):
        _lib.check(self.lib.stb_set_batch(self.handle, int(n_sys)))
        self.n_sys = int(n_sys)
        self.active_params = None

    def _blk(self, blk, n):
        if n == 0:
            return None
        if blk is None:
            return np.zeros((n, n))
        blk = _lib.f64(blk)
        if blk.shape != (n, n):
            raise ValueError('block must be %dx%d' % (n, n))
        return blk

    def set_decay_blocks(self, blk_S, blk_A, system: int = -1):
        bS, bA = self._blk(blk_S, self.sS), self._blk(blk_A, self.sA)
        self.active_params = None
        _lib.check(self.lib.stb_set_blocks(self.handle, system, 0, _lib.ptr(bS), _lib.ptr(bA)))

    def set_abs_blocks(self, blk_S, blk_A, system: int = -1):
        bS, bA = self._blk(blk_S, self.sS), self._blk(blk_A, self.sA)
        self.active_params = None
        _lib.check(self.lib.stb_set_blocks(self.handle, system, 1, _lib.ptr(bS), _lib.ptr(bA)))

    def set_strengths(self, strengths: Sequence[float], system: int = -1):
        s = _lib.f64(strengths)
        if s.shape != (self.n_proc,):
            raise ValueError('need %d strengths' % self.n_proc)
        self.active_params = None
        rc = self.lib.stb_set_strengths(self.handle, system, _lib.ptr(s))
        if rc != 0:
            raise LatticeError(_lib.last_error())

    # ------------------------------------------------------------------ reference-layout view
    def num_interactions(self, system: int = 0) -> int:
        m = self.lib.stb_num_interactions(self.handle, system)
        if m < 0:
            raise _lib.StbError(-1, _lib.last_error())
        return int(m)

    def export_interactions(self, system: int = 0, jac: bool = True):
        """`(N_indices [M,2] u32, et_rows [M,4] i32, et_w [M] f64, jac_indices [2M,3] u32)` in the
        reference's interaction order (`precalculate.py:349-392, 603-623`)."""
        M = self.num_interactions(system)
        N_idx = np.zeros((M, 2), dtype=np.uint32)
        rows = np.zeros((M, 4), dtype=np.int32)
        w = np.zeros(M, dtype=np.float64)
        jac_idx = np.zeros((2 * M, 3), dtype=np.uint32) if jac else None
        if M:
            _lib.check(self.lib.stb_export_interactions(
                self.handle, system, _lib.ptr(N_idx), _lib.ptr(rows), _lib.ptr(w), _lib.ptr(jac_idx)))
        return N_idx, rows, w, jac_idx

    def ET_matrix(self, system: int = 0):
        """The reference's `ET_matrix [N, M]` CSR (`precalculate.py:410-412`) and `N_indices`,
        `jac_indices`, assembled from the device export."""
        N_idx, rows, w, jac_idx = self.export_interactions(system)
        M = len(w)
        if M == 0:
            return (sp.csr_matrix((self.N, 0), dtype=np.float64),
                    np.zeros((0, 2), dtype=np.uint32), np.zeros((0, 3), dtype=np.uint32))
        vals = np.column_stack((-w, w, -w, w)).ravel()
        cols = np.repeat(np.arange(M, dtype=np.int64), 4)
        ET = sp.csr_matrix((vals, (rows.ravel().astype(np.int64), cols)), shape=(self.N, M),
                           dtype=np.float64)
        return ET, N_idx, jac_idx

    # ------------------------------------------------------------------ operators
    def _vecs(self, y, first, n):
        y = _lib.f64(y)
        if y.ndim == 1:
            y = y[None, :]
        if n is None:
            n = y.shape[0]
        if y.shape != (n, self.N):
            raise ValueError('state array must be [%d, %d]' % (n, self.N))
        return y, n

    def rhs(self, y, pulse: bool, first: int = 0):
        """`_rate_eq_pulse` (pulse) / `_rate_eq` for systems `first..first+len(y)`."""
        single = np.ndim(y) == 1
        y, n = self._vecs(y, first, None)
        out = np.empty_like(y)
        _lib.check(self.lib.stb_rhs(self.handle, first, n, int(bool(pulse)), _lib.ptr(y), _lib.ptr(out)))
        return out[0] if single else out

    def jv(self, y, v, pulse: bool, first: int = 0):
        """`_jac_rate_eq[_pulse](y) @ v` without the dense Jacobian."""
        single = np.ndim(y) == 1
        y, n = self._vecs(y, first, None)
        v, _ = self._vecs(v, first, n)
        out = np.empty_like(y)
        _lib.check(self.lib.stb_jv(self.handle, first, n, int(bool(pulse)), _lib.ptr(y), _lib.ptr(v),
                                   _lib.ptr(out)))
        return out[0] if single else out

    def integrate(self, t, y0, pulse: bool, rtol: float = 1e-3, atol: float = 1e-15,
                  nsteps: int = 1000, first: int = 0, want_y: bool = True, want_avg: bool = False):
        """Integrate systems `first..first+n` over the output grid `t` (shared).  Returns
        `(y_out [n, n_t, N] | None, avg [n, n_t, sS+sA] | None, stats list)`."""
        t = _lib.f64(t).ravel()
        y0, n = self._vecs(y0, first, None)
        n_t = len(t)
        y_out = np.zeros((n, n_t, self.N)) if want_y else None
        avg = np.zeros((n, n_t, self.sS + self.sA)) if want_avg else None
        stats = (_lib.StbStats * n)()
        _lib.check(self.lib.stb_integrate(
            self.handle, first, n, int(bool(pulse)), n_t, _lib.ptr(t), _lib.ptr(y0), float(rtol),
            float(atol), int(nsteps), _lib.ptr(y_out), _lib.ptr(avg), C.cast(stats, C.c_void_p)))
        return y_out, avg, [s.as_dict() for s in stats]

    def integrate_dynamics(self, t_pulse, t_sol, y0, rtol: float = 1e-3, atol: float = 1e-15,
                           nsteps: int = 1000, first: int = 0, want_y: bool = True,
                           want_avg: bool = True, want_pulse: bool = False, cumulative: bool = False):
        """Pulse then relaxation of systems `first..first+n` in one C-ABI call
        (`stb_integrate_dynamics`): the relaxation starts on the device from the last pulse row.
        Returns `(y_pulse [n, n_tp, N] | None, y [n, n_t, N] | None, avg [n, n_t, sS+sA] | None,
        stats_pulse, stats_relax)`."""
        t_pulse = _lib.f64(t_pulse).ravel()
        t_sol = _lib.f64(t_sol).ravel()
        y0, n = self._vecs(y0, first, None)
        y_p = np.empty((n, len(t_pulse), self.N)) if want_pulse else None
        y = np.empty((n, len(t_sol), self.N)) if want_y else None
        avg = np.empty((n, len(t_sol), self.sS + self.sA)) if want_avg else None
        st_p = (_lib.StbStats * n)()
        st_r = (_lib.StbStats * n)()
        # cumulative: `avg` is the running time integral of the emission curves (pulsed steady state)
        self.set_option('avg_cumulative', 1.0 if cumulative else 0.0)
        _lib.check(self.lib.stb_integrate_dynamics(
            self.handle, first, n, len(t_pulse), _lib.ptr(t_pulse), len(t_sol), _lib.ptr(t_sol),
            _lib.ptr(y0), float(rtol), float(atol), int(nsteps), int(nsteps), _lib.ptr(y_p),
            _lib.ptr(y), _lib.ptr(avg), C.cast(st_p, C.c_void_p), C.cast(st_r, C.c_void_p)))
        return y_p, y, avg, [s.as_dict() for s in st_p], [s.as_dict() for s in st_r]

    def integrate_dynamics_dev(self, t_pulse_ptr: int, n_tp: int, t_sol_ptr: int, n_t: int,
                               y0_planes_ptr: int, n: int, rtol: float, atol: float,
                               nsteps: int = 1000, first: int = 0, y_pulse_ptr: int = 0,
                               y_out_ptr: int = 0, avg_out_ptr: int = 0):
        """`stb_integrate_dynamics_dev`: pulse + relaxation with every operand a raw device pointer;
        only the per-system statistics come back to the host."""
        st_p = (_lib.StbStats * n)()
        st_r = (_lib.StbStats * n)()
        _lib.check(self.lib.stb_integrate_dynamics_dev(
            self.handle, first, n, n_tp, C.c_void_p(t_pulse_ptr), n_t, C.c_void_p(t_sol_ptr),
            C.c_void_p(y0_planes_ptr), float(rtol), float(atol), int(nsteps), int(nsteps),
            C.c_void_p(y_pulse_ptr or None), C.c_void_p(y_out_ptr or None),
            C.c_void_p(avg_out_ptr or None), C.cast(st_p, C.c_void_p), C.cast(st_r, C.c_void_p)))
        return [s.as_dict() for s in st_p], [s.as_dict() for s in st_r]

    def last_timing_dynamics(self):
        """`(pulse ms, relaxation ms)` of the kernels of the last `integrate_dynamics[_dev]` call."""
        a, b = C.c_float(), C.c_float()
        _lib.check(self.lib.stb_last_timing_dynamics(self.handle, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def integrate_dev(self, t_dev_ptr: int, n_t: int, y0_planes_ptr: int, n: int, pulse: bool,
                      rtol: float, atol: float, nsteps: int = 1000, first: int = 0,
                      y_out_ptr: int = 0, avg_out_ptr: int = 0):
        """`stb_integrate_dev`: every operand is a raw device pointer (e.g. `tensor.data_ptr()`),
        nothing is copied between host and device except the per-system statistics."""
        stats = (_lib.StbStats * n)()
        _lib.check(self.lib.stb_integrate_dev(
            self.handle, first, n, int(bool(pulse)), n_t, C.c_void_p(t_dev_ptr),
            C.c_void_p(y0_planes_ptr), float(rtol), float(atol), int(nsteps),
            C.c_void_p(y_out_ptr or None), C.c_void_p(avg_out_ptr or None),
            C.cast(stats, C.c_void_p)))
        return [s.as_dict() for s in stats]

    def to_planes(self, y_host, out_ptr: int):
        """Reference-layout host vectors `[n, N]` -> plane layout in the device buffer `out_ptr`."""
        y, n = self._vecs(y_host, 0, None)
        _lib.check(self.lib.stb_to_planes(self.handle, n, _lib.ptr(y), C.c_void_p(out_ptr)))

    @property
    def stream(self) -> int:
        """`cudaStream_t` (as int) all kernels of this system run on."""
        return int(self.lib.stb_stream(self.handle) or 0)

    def last_timing(self):
        ms = C.c_float()
        launches = C.c_int64()
        _lib.check(self.lib.stb_last_timing(self.handle, C.byref(ms), C.byref(launches)))
        return float(ms.value), int(launches.value)

    def last_profile(self):
        """Per-phase clock cycles of the last integrate call (development aid)."""
        out = (C.c_double * 10)()
        _lib.check(self.lib.stb_last_profile(self.handle, out))
        names = ['rhs', 'arnoldi', 'dots', 'barrier', 'vector', 'givens', 'arnoldi_stream', 'rest', 'G', 'groups']
        return dict(zip(names, [float(v) for v in out]))
