"""CPU restatement (numpy/scipy) of the reference's microscopic rate-equation path.

TEST INFRASTRUCTURE - this is the checker, never the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import it; `simetuc_b200/` never does.

Parity status: PINNED.  Every function below is checked (tests/test_oracle.py) against
  * the reference's own hand-checked known-answer matrices (`test/test_setup/test_setup.py:25-598`),
  * the golden trajectory `test/test_simulations/{t,y}_sol_2S_2A.hdf5`,
  * outputs of the unmodified reference functions run in the build container
    (`oracle/make_golden.py` -> `tests/golden/*.npz`), and live against
    `/root/reference` when that tree is present.

The integrator arithmetic of the reference lives in a third-party dependency,
`scipy.integrate.ode('vode', method='bdf')` (scipy unpinned: `requirements.txt:1`
`scipy >=0.18.1`; call site `odesolver.py:100-104,126`).  The oracle drives the same
callbacks with scipy's `lsoda` (north_star: "replacing scipy's LSODA/BDF"), because this
image's scipy 1.18.1 `vode` mis-routes `jac_params` and is >20x slower.

Each function cites the reference lines it follows.  Ion types: 0 = sensitizer (S),
1 = activator (A).  State vector layout: ion-major, ions in lattice order, ion q's
states at `y[gs_q : gs_q + s_type(q)]` (`lattice.py:279-292`).
"""
import numpy as np
import scipy.sparse as sp
from scipy.integrate import ode as _scipy_ode


def _code(ion):
    return int(getattr(ion, 'value', ion))


# ----------------------------------------------------------------------------- lattice side
def ground_states(ion_type, n_S_states, n_A_states):
    """GS index per ion and initial populations.  Follows `lattice.py:259-298`.

    Returns `(index_S_i, index_A_j, initial_population)`: int64 arrays of length n_ions with
    -1 where the ion is not of that type (or that type has no states), and a uint64 vector
    with 1 at every ground state."""
    ion_type = np.asarray(ion_type)
    is_A = ion_type != 0
    width = np.where(is_A, n_A_states, n_S_states).astype(np.int64)
    start = np.concatenate(([0], np.cumsum(width)[:-1])) if len(width) else np.zeros(0, np.int64)
    index_S_i = np.where(~is_A & (n_S_states != 0), start, -1).astype(np.int64)
    index_A_j = np.where(is_A & (n_A_states != 0), start, -1).astype(np.int64)
    pop = np.zeros(int(width.sum()), dtype=np.uint64)
    pop[index_S_i[index_S_i >= 0]] = 1
    pop[index_A_j[index_A_j >= 0]] = 1
    return index_S_i, index_A_j, pop


def interaction_lists(ion_type, dist_array, index_S_i, index_A_j, n_S_states, n_A_states):
    """All-to-all partner lists.  Follows `lattice.py:301-374`: for every ion, the GS index and
    the distance of every other ion (those with non-zero distance, ascending ion order), split
    by partner type.  Returns 8 ragged lists
    `(index_S_k, index_S_l, index_A_k, index_A_l, dist_S_k, dist_S_l, dist_A_k, dist_A_l)`;
    `X_k` = S partners, `X_l` = A partners, one entry per S ion (`S_*`) or A ion (`A_*`)."""
    ion_type = np.asarray(ion_type)
    dist_array = np.asarray(dist_array, dtype=np.float64)
    out = {k: [] for k in ('iSk', 'iSl', 'iAk', 'iAl', 'dSk', 'dSl', 'dAk', 'dAl')}
    for q in range(len(ion_type)):
        row = dist_array[q]
        others = np.flatnonzero(row)
        partner_is_A = ion_type[others] != 0
        to_A, to_S = others[partner_is_A], others[~partner_is_A]
        if ion_type[q] != 0:
            if n_A_states == 0:
                continue
            tag = 'A'
        else:
            if n_S_states == 0:
                continue
            tag = 'S'
        out['i%sl' % tag].append(index_A_j[to_A].astype(np.int64))
        out['d%sl' % tag].append(row[to_A])
        out['i%sk' % tag].append(index_S_i[to_S].astype(np.int64))
        out['d%sk' % tag].append(row[to_S])
    return (out['iSk'], out['iSl'], out['iAk'], out['iAl'],
            out['dSk'], out['dSl'], out['dAk'], out['dAl'])


# ----------------------------------------------------------------------------- per-type blocks
def absorption_blocks(n_S_states, n_A_states, excitations):
    """Per-type absorption blocks summed over every process of every ACTIVE excitation.
    Follows `precalculate.py:79-127`: with `P = pump_rate*power_dens`, `g = degeneracy`:
    `B[i,i] -= P; B[i,f] += g*P; B[f,i] += P; B[f,f] -= g*P` on the block of the excited ion
    type; a process whose states do not exist is skipped (`:107,115`)."""
    blk = [np.zeros((n_S_states, n_S_states)), np.zeros((n_A_states, n_A_states))]
    for exc_list in excitations.values():
        for exc in exc_list:
            if exc.active is False:
                continue
            t = _code(exc.transition.ion)
            n = (n_S_states, n_A_states)[t]
            i, f = exc.transition.state_i, exc.transition.state_f
            if n == 0 or not (i < n and f < n):
                continue
            one = np.zeros((n, n))
            one[i, i] = -exc.pump_rate
            one[i, f] = +exc.degeneracy * exc.pump_rate
            one[f, i] = +exc.pump_rate
            one[f, f] = -exc.degeneracy * exc.pump_rate
            blk[t] = blk[t] + one * exc.power_dens
    return blk[0], blk[1]


def decay_blocks(n_S_states, n_A_states, decay):
    """Per-type decay blocks `D = (B^T - I) * k` (columns scaled by the decay rates), row 0
    replaced by minus the column sums.  Follows `precalculate.py:130-213`: branching ratios to
    the ground state are discarded (`:153-156`), transitions that name non-existent states are
    skipped (`:142,146` - the A check tests `state_i` twice, mirrored here), decay rates of
    non-existent states are skipped (`:172,175`)."""
    out = []
    for t, n, bkey, kkey in ((0, n_S_states, 'branching_S', 'decay_S'),
                             (1, n_A_states, 'branching_A', 'decay_A')):
        if n == 0:
            out.append(None)
            continue
        B = np.zeros((n, n))
        for tr in decay[bkey]:
            ok = (tr.state_i < n and tr.state_f < n) if t == 0 else (tr.state_i < n)
            if ok:
                B[tr.state_i, tr.state_f] = tr.branching_ratio
        B[:, 0] = 0
        k = np.zeros(n)
        for tr in decay[kkey]:
            if tr.state_i < n:
                k[tr.state_i] = tr.decay_rate
        D = (B.T - np.eye(n)) * k[None, :]
        D[0, :] = -D.sum(axis=0)
        out.append(D)
    return out[0], out[1]


def block_diagonal(block_S, block_A, index_S_i, index_A_j):
    """Ion-ordered block-diagonal CSR operator (`precalculate.py:51-76, 217-228`)."""
    blocks = []
    for s, a in zip(index_S_i, index_A_j):
        if s != -1:
            if block_S is not None and block_S.size:
                blocks.append(block_S)
        elif a != -1:
            if block_A is not None and block_A.size:
                blocks.append(block_A)
    if not blocks:
        return sp.csr_matrix((0, 0), dtype=np.float64)
    return sp.block_diag([sp.csr_matrix(b) for b in blocks], format='csr', dtype=np.float64)


def total_absorption_matrix(n_S_states, n_A_states, excitations, index_S_i, index_A_j):
    bS, bA = absorption_blocks(n_S_states, n_A_states, excitations)
    return block_diagonal(bS, bA, index_S_i, index_A_j)


def decay_matrix(n_S_states, n_A_states, decay, index_S_i, index_A_j):
    bS, bA = decay_blocks(n_S_states, n_A_states, decay)
    return block_diagonal(bS, bA, index_S_i, index_A_j)


# ----------------------------------------------------------------------------- pair ET
def _active_pair_processes(energy_transfer):
    """Pair processes with non-zero strength, in dict order, as
    `(donor_type, partner_type, ii, fi, if, ff, mult, strength, name)`
    (`precalculate.py:353-354`: `not np.isclose(strength, 0)`; `util.py:247-249`)."""
    procs = []
    for name, p in energy_transfer.items():
        if len(p.type) != 2 or np.isclose(p.strength, 0.0):
            continue
        ii, fi, if_, ff = p.indices
        procs.append((_code(p.type[0]), _code(p.type[1]), ii, fi, if_, ff,
                      int(p.mult), float(p.strength), name))
    return procs


def check_process_states(energy_transfer, n_S_states, n_A_states):
    """State-count validation of `precalculate.py:325-345` (it uses `>` not `>=`, and the S-A
    branch tests the same index slice against both counts - both quirks mirrored).  Returns the
    name of the first offending process or None."""
    for dt, pt, ii, fi, if_, ff, _m, _s, name in _active_pair_processes(energy_transfer):
        idx = np.array([ii, fi, if_, ff])
        nS, nA = n_S_states, n_A_states
        bad = False
        if (dt, pt) == (1, 1):
            bad = np.any(idx > nA)
        elif (dt, pt) == (1, 0):
            bad = np.any(idx[::2] > nA) or np.any(idx[1::2] > nS)
        elif (dt, pt) == (0, 0):
            bad = np.any(idx > nS)
        elif (dt, pt) == (0, 1):
            bad = np.any(idx[::2] > nA) or np.any(idx[0::2] > nS)
        if bad:
            return name
    return None


def ET_matrices(index_S_i, index_A_j, energy_transfer,
                index_S_k, index_S_l, index_A_k, index_A_l,
                dist_S_k, dist_S_l, dist_A_k, dist_A_l, n_S_states, n_A_states):
    """ET operator `[N, M]` (4 entries per interaction column: -w,+w on the donor ion's initial /
    final state, -w,+w on the partner's) and `N_indices [M,2] uint32` = the two initial states
    whose product drives interaction r.  `w = strength * d**(-mult)`.  Follows
    `precalculate.py:235-415`.  Interaction order (`:349-392`): ions in lattice order; for an A
    ion every active A-A process (dict order) over all A partners, then every A-S process over
    all S partners; for an S ion the S-S processes then the S-A ones."""
    index_S_i = np.asarray(index_S_i).ravel()
    index_A_j = np.asarray(index_A_j).ravel()
    nS = int(np.count_nonzero(index_S_i != -1))
    nA = int(np.count_nonzero(index_A_j != -1))
    N = n_S_states * nS + n_A_states * nA
    empty = (sp.csr_matrix((N, 0), dtype=np.float64), np.zeros((0, 2), dtype=np.uint32))
    if N <= 2:
        return empty
    procs = _active_pair_processes(energy_transfer)
    rows, vals, n_i, n_j = [], [], [], []
    cS = cA = 0
    for q in range(nS + nA):
        if index_A_j[q] != -1 and n_A_states != 0:
            gs, mine = int(index_A_j[q]), 1
            lists = {1: (index_A_l[cA], dist_A_l[cA]), 0: (index_A_k[cA], dist_A_k[cA])}
            order = (1, 0)
            cA += 1
        elif index_S_i[q] != -1 and n_S_states != 0:
            gs, mine = int(index_S_i[q]), 0
            lists = {0: (index_S_k[cS], dist_S_k[cS]), 1: (index_S_l[cS], dist_S_l[cS])}
            order = (0, 1)
            cS += 1
        else:
            continue
        for partner in order:
            idx, dist = lists[partner]
            idx = np.asarray(idx, dtype=np.int64).ravel()
            dist = np.asarray(dist, dtype=np.float64).ravel()
            for dt, pt, ii, fi, if_, ff, mult, strength, _n in procs:
                if dt != mine or pt != partner:
                    continue
                w = dist ** (-mult) * strength
                m = len(idx)
                r4 = np.empty((m, 4), dtype=np.int64)
                r4[:, 0] = gs + ii
                r4[:, 1] = gs + if_
                r4[:, 2] = idx + fi
                r4[:, 3] = idx + ff
                rows.append(r4)
                vals.append(np.column_stack((-w, w, -w, w)))
                n_i.append(np.full(m, gs + ii, dtype=np.uint32))
                n_j.append((idx + fi).astype(np.uint32))
    M = sum(len(a) for a in n_i)
    if M == 0:
        return empty
    rows = np.concatenate(rows).ravel()
    vals = np.concatenate(vals).ravel()
    cols = np.repeat(np.arange(M, dtype=np.int64), 4)
    ET = sp.csr_matrix((vals, (rows, cols)), shape=(N, M), dtype=np.float64)
    N_indices = np.column_stack((np.concatenate(n_i), np.concatenate(n_j))).astype(np.uint32)
    return ET, N_indices


def jac_indices(N_indices):
    """`[2M,3]`: rows `(r, N_i[r], N_j[r])` and `(r, N_j[r], N_i[r])` interleaved.
    Follows `precalculate.py:603-623`."""
    N_indices = np.asarray(N_indices)
    M = N_indices.shape[0]
    out = np.empty((2 * M, 3), dtype=np.uint32)
    out[0::2, 0] = out[1::2, 0] = np.arange(M, dtype=np.uint32)
    out[0::2, 1] = N_indices[:, 0]
    out[1::2, 1] = N_indices[:, 1]
    out[0::2, 2] = N_indices[:, 1]
    out[1::2, 2] = N_indices[:, 0]
    return out


def empty_coop(N):
    """Empty cooperative operands accepted by the solver entry points."""
    return (sp.csr_matrix((N, 0), dtype=np.float64), np.zeros((0, 3), dtype=np.uint32),
            np.zeros((0, 4), dtype=np.uint32))


# ----------------------------------------------------------------------------- RHS / Jacobian
def rate_eq(t, y, decay_matrix, UC_matrix, N_indices, coop_ET_matrix, coop_N_indices):
    """`dy = D y + ET (y[Ni]*y[Nj]) + ETc (y[c0]*y[c1]*y[c2])`.  Follows `odesolver.py:33-43`."""
    pair = y[N_indices[:, 0]] * y[N_indices[:, 1]]
    out = decay_matrix.dot(y) + UC_matrix.dot(pair)
    if coop_N_indices.shape[0]:
        c = coop_N_indices
        out = out + coop_ET_matrix.dot(y[c[:, 0]] * y[c[:, 1]] * y[c[:, 2]])
    return out


def rate_eq_pulse(t, y, abs_matrix, decay_matrix, UC_matrix, N_indices,
                  coop_ET_matrix, coop_N_indices):
    """`abs y + rate_eq`.  Follows `odesolver.py:65-71`."""
    return abs_matrix.dot(y) + rate_eq(t, y, decay_matrix, UC_matrix, N_indices,
                                       coop_ET_matrix, coop_N_indices)


def jac_rate_eq(t, y, decay_matrix, UC_matrix, jac_idx, coop_ET_matrix, coop_jac_indices):
    """Dense Jacobian `D + ET nJ (+ coop)`, `nJ[r, col] = y[other]`.  Follows `odesolver.py:46-62`."""
    M, N = UC_matrix.shape[1], UC_matrix.shape[0]
    nJ = sp.csr_matrix((y[jac_idx[:, 2]], (jac_idx[:, 0], jac_idx[:, 1])), shape=(M, N))
    J = decay_matrix.toarray() + UC_matrix.dot(nJ).toarray()
    if coop_jac_indices.shape[0]:
        c = coop_jac_indices
        nJc = sp.csr_matrix((y[c[:, 2]] * y[c[:, 3]], (c[:, 0], c[:, 1])),
                            shape=(coop_ET_matrix.shape[1], N))
        J = J + coop_ET_matrix.dot(nJc).toarray()
    return J


def jac_rate_eq_pulse(t, y, abs_matrix, decay_matrix, UC_matrix, jac_idx,
                      coop_ET_matrix, coop_jac_indices):
    """Follows `odesolver.py:74-80`."""
    return abs_matrix.toarray() + jac_rate_eq(t, y, decay_matrix, UC_matrix, jac_idx,
                                              coop_ET_matrix, coop_jac_indices)


def jac_vec(y, v, lin_matrix, UC_matrix, N_indices):
    """`J(y) v = L v + ET (y[Nj] v[Ni] + y[Ni] v[Nj])` without forming J."""
    ni, nj = N_indices[:, 0], N_indices[:, 1]
    return lin_matrix.dot(v) + UC_matrix.dot(y[nj] * v[ni] + y[ni] * v[nj])


# ----------------------------------------------------------------------------- integration
def solve_ode(t_arr, fun, jfun, y0, rtol=1e-3, atol=1e-15, nsteps=1000,
              integrator='lsoda', **kw):
    """`y_arr[0] = y0; y_arr[k] = integrate(t_arr[k])` with one continuing integrator object.
    Follows `odesolver.py:85-137` (stepping loop `:123-127`); `integrator='vode'` reproduces the
    reference call exactly (`method='bdf'`), `'lsoda'` is the oracle default (module docstring).
    `fun(t, y)` / `jfun(t, y)` are closures (argument binding of `:103-104` done by the caller)."""
    y0 = np.asarray(y0, dtype=np.float64)
    out = np.zeros((len(t_arr), len(y0)))
    solver = _scipy_ode(fun, jfun)
    if integrator == 'vode':
        solver.set_integrator('vode', rtol=rtol, atol=atol, method='bdf', nsteps=nsteps, **kw)
    else:
        solver.set_integrator('lsoda', rtol=rtol, atol=atol, nsteps=nsteps, **kw)
    solver.set_initial_value(y0, t_arr[0])
    out[0] = y0
    k = 1
    while solver.successful() and k < len(t_arr):
        out[k] = solver.integrate(t_arr[k])
        k += 1
    return out


def solve_pulse(t_pulse, y0, abs_matrix, decay_matrix, UC_matrix, N_indices, jac_idx,
                coop=None, **kw):
    """Follows `odesolver.py:178-193`."""
    cM, cN, cJ = coop if coop is not None else empty_coop(len(y0))
    return solve_ode(
        t_pulse,
        lambda t, y: rate_eq_pulse(t, y, abs_matrix, decay_matrix, UC_matrix, N_indices, cM, cN),
        lambda t, y: jac_rate_eq_pulse(t, y, abs_matrix, decay_matrix, UC_matrix, jac_idx, cM, cJ),
        y0, **kw)


def solve_relax(t_sol, y0, decay_matrix, UC_matrix, N_indices, jac_idx, coop=None, **kw):
    """Follows `odesolver.py:196-209`."""
    cM, cN, cJ = coop if coop is not None else empty_coop(len(y0))
    return solve_ode(
        t_sol,
        lambda t, y: rate_eq(t, y, decay_matrix, UC_matrix, N_indices, cM, cN),
        lambda t, y: jac_rate_eq(t, y, decay_matrix, UC_matrix, jac_idx, cM, cJ),
        y0, **kw)


# ----------------------------------------------------------------------------- results
def lifetimes(decay):
    """`1/k` of every decay transition (`precalculate.py:702-709`)."""
    return [1.0 / float(tr.decay_rate) for tr in list(decay['decay_S']) + list(decay['decay_A'])]


def t_simulation(decay):
    """`round(15 * max lifetime, 8)` (`simulations.py:1092-1093`)."""
    return float(np.round(15 * np.max(lifetimes(decay)), 8))


def time_grids(t_pulse_width, tf, n_steps):
    """Pulse grid `linspace(0, t_pulse, 2)` and relaxation grid
    `logspace(log10 t_pulse, log10 tf, N)` (`simulations.py:1125-1156`)."""
    return (np.linspace(0, t_pulse_width, 2, dtype=np.float64),
            np.logspace(np.log10(t_pulse_width), np.log10(tf), n_steps, dtype=np.float64))


def avg_populations(y_sol, index_S_i, index_A_j, n_S_states, n_A_states):
    """Per state kind, mean over the ions of that type, clipped at 0; S states then A states.
    Follows `simulations.py:100-128`.  Returns `[sS+sA, N_t]`."""
    y_sol = np.asarray(y_sol)
    index_S_i = np.asarray(index_S_i).ravel()
    index_A_j = np.asarray(index_A_j).ravel()
    out = []
    for idx, n in ((index_S_i, n_S_states), (index_A_j, n_A_states)):
        gs = idx[idx != -1]
        for s in range(n):
            if len(gs):
                out.append(np.clip(y_sol[:, gs + s].sum(axis=1) / len(gs), 0, None))
            else:
                out.append(np.zeros(y_sol.shape[0]))
    return np.array(out)


def final_populations(avg, n_S_states, n_A_states, S_conc, A_conc):
    """Last point of every average curve times the ion concentration (`simulations.py:302-312`)."""
    conc = np.array([S_conc] * n_S_states + [A_conc] * n_A_states, dtype=np.float64)
    return np.array([c[-1] for c in avg]) * conc


def cumulative_steady_state(t_sol, y_sol):
    """Cumulative trapezoid of every state over `t_sol`, first row 0 (`simulations.py:579-591`)."""
    dt = np.diff(t_sol)[:, None]
    out = np.zeros_like(y_sol)
    out[1:] = np.cumsum(0.5 * (y_sol[1:] + y_sol[:-1]) * dt, axis=0)
    return out


# ----------------------------------------------------------------------------- cooperative ET
def coop_ET_matrices(index_S_i, index_A_j, energy_transfer, index_A_k, dist_A_k, dist_S_k,
                     n_S_states, n_A_states, d_max_coop=np.inf):
    """Cooperative (S,S,A) three-body operator `[N, Mc]` (6 entries per column) and
    `coop_N_indices [Mc,3] uint32`.  Follows `precalculate.py:427-599`: the first process of
    type (S,S,A) with non-zero strength (`:515-520`); for every A ion l (A order) every
    unordered pair (i<k) of S ions with `max(d_li, d_lk) < d_max_coop` (`:471-480`);
    `w = C*(1/(d_li d_lk)^m + 1/(d_li d_ik)^m + 1/(d_ik d_lk)^m)` (`:499-506,585`)."""
    index_S_i = np.asarray(index_S_i).ravel()
    index_A_j = np.asarray(index_A_j).ravel()
    gs_S = index_S_i[index_S_i != -1]
    gs_A = index_A_j[index_A_j != -1]
    N = n_S_states * len(gs_S) + n_A_states * len(gs_A)
    empty = (sp.csr_matrix((N, 0), dtype=np.float64), np.zeros((0, 3), dtype=np.uint32))
    proc = None
    for p in energy_transfer.values():
        if len(p.type) == 3 and not np.isclose(p.strength, 0.0) and \
                tuple(_code(t) for t in p.type) == (0, 0, 1):
            proc = p
            break
    n_dSk = sum(np.size(r) for r in dist_S_k)
    n_dAk = sum(np.size(r) for r in dist_A_k)
    if N <= 5 or proc is None or n_dSk == 0 or n_dAk == 0:
        return empty
    ii, ki, li, if_, kf, lf = proc.indices
    mult, C = int(proc.mult), float(proc.strength)
    pos_of_gs = {int(g): n for n, g in enumerate(gs_S)}
    cols = []
    for l, gl in enumerate(gs_A):
        idx = np.asarray(index_A_k[l], dtype=np.int64).ravel()
        d = np.asarray(dist_A_k[l], dtype=np.float64).ravel()
        a, b = np.triu_indices(len(idx), k=1)          # combinations order (a<b, a-major)
        keep = np.maximum(d[a], d[b]) < d_max_coop
        a, b = a[keep], b[keep]
        for gi, gk, dli, dlk in zip(idx[a], idx[b], d[a], d[b]):
            ni, nk = pos_of_gs[int(gi)], pos_of_gs[int(gk)]
            dik = float(np.asarray(dist_S_k[ni]).ravel()[nk - 1])
            cols.append((int(gi), int(gk), int(gl), dli, dlk, dik))
    Mc = len(cols)
    if Mc == 0:
        return empty
    arr = np.array(cols, dtype=np.float64)
    gi, gk, gl = (arr[:, c].astype(np.int64) for c in range(3))
    dli, dlk, dik = arr[:, 3], arr[:, 4], arr[:, 5]
    w = C * (1 / (dli * dlk) ** mult + 1 / (dli * dik) ** mult + 1 / (dik * dlk) ** mult)
    rows = np.column_stack((gi + ii, gi + if_, gk + ki, gk + kf, gl + li, gl + lf)).ravel()
    vals = np.column_stack((-w, w, -w, w, -w, w)).ravel()
    cidx = np.repeat(np.arange(Mc, dtype=np.int64), 6)
    M = sp.csr_matrix((vals, (rows, cidx)), shape=(N, Mc), dtype=np.float64)
    Nidx = np.column_stack((gi + ii, gk + ki, gl + li)).astype(np.uint32)
    return M, Nidx


def coop_jac_indices(coop_N_indices):
    """`[3Mc,4]` rows `(r, col, ya, yb)`: `(r, l, i, k), (r, k, i, l), (r, i, k, l)`.
    Follows `precalculate.py:626-699`."""
    c = np.asarray(coop_N_indices)
    Mc = c.shape[0]
    out = np.empty((3 * Mc, 4), dtype=np.uint32)
    out[:, 0] = np.repeat(np.arange(Mc, dtype=np.uint32), 3)
    if Mc:
        out[0::3, 1], out[1::3, 1], out[2::3, 1] = c[:, 2], c[:, 1], c[:, 0]
        out[0::3, 2], out[1::3, 2], out[2::3, 2] = c[:, 0], c[:, 0], c[:, 1]
        out[0::3, 3], out[1::3, 3], out[2::3, 3] = c[:, 1], c[:, 2], c[:, 2]
    return out


# ----------------------------------------------------------------------------- row-sampled RHS / Jv
def _mic_row(p, positions, cell):
    """Minimum-image distances of one site to all sites of the periodic supercell `cell` (rows =
    supercell vectors; None = open boundaries): the distance ASE's `get_all_distances(mic=True)`
    returns at `lattice.py:239`.  Nearest fractional image, then the 27 neighbouring images."""
    d = np.asarray(positions, dtype=np.float64) - np.asarray(p, dtype=np.float64)[None, :]
    if cell is None:
        return np.sqrt((d * d).sum(-1))
    cell = np.asarray(cell, dtype=np.float64)
    frac = d @ np.linalg.inv(cell)
    d = (frac - np.rint(frac)) @ cell
    best = np.full(len(d), np.inf)
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                e = d + (i * cell[0] + j * cell[1] + k * cell[2])[None, :]
                np.minimum(best, (e * e).sum(-1), out=best)
    return np.sqrt(best)


def _ion_rows(q, ion_type, gs, positions, cell, lin_blocks, procs, y, v=None):
    """Rows of ion q of `rate_eq[_pulse]` (and of `J(y) v` when v is given), O(n_ions) work."""
    T = int(ion_type[q] != 0)
    s = lin_blocks[T].shape[0]
    d = _mic_row(positions[q], positions, cell)
    others = np.flatnonzero(d)                               # lattice.py:335: self excluded by distance 0
    yq = y[gs[q]:gs[q] + s]
    out = lin_blocks[T].dot(yq)
    outv = None
    if v is not None:
        vq = v[gs[q]:gs[q] + s]
        outv = lin_blocks[T].dot(vq)
    for dt, pt, ii, fi, if_, ff, mult, strength, _n in procs:
        # ion q as the donor of the process (precalculate.py:277-291: -w at gs_i+ii, +w at gs_i+if)
        if dt == T:
            k = others[(ion_type[others] != 0) == bool(pt)]
            w = strength * d[k] ** (-mult)
            rate = np.sum(w * yq[ii] * y[gs[k] + fi])
            out[ii] -= rate
            out[if_] += rate
            if v is not None:
                dr = np.sum(w * (vq[ii] * y[gs[k] + fi] + yq[ii] * v[gs[k] + fi]))
                outv[ii] -= dr
                outv[if_] += dr
        # ion q as the partner (-w at gs_k+fi, +w at gs_k+ff)
        if pt == T:
            i = others[(ion_type[others] != 0) == bool(dt)]
            w = strength * d[i] ** (-mult)
            rate = np.sum(w * y[gs[i] + ii] * yq[fi])
            out[fi] -= rate
            out[ff] += rate
            if v is not None:
                dr = np.sum(w * (v[gs[i] + ii] * yq[fi] + y[gs[i] + ii] * vq[fi]))
                outv[fi] -= dr
                outv[ff] += dr
    return out, outv


def rate_eq_rows(ions, ion_type, positions, cell, n_S_states, n_A_states, energy_transfer,
                 lin_S, lin_A, y, v=None):
    """The rows of `rate_eq[_pulse](y)` - and of `jac_rate_eq[_pulse](y) @ v` - that belong to the
    ions in `ions`, computed ion by ion from positions (no interaction list, no ET matrix): the
    check for lattices whose reference operands (56 B per interaction, dense Jacobian) do not fit
    in memory.  Regroups `odesolver.py:33-43` / `:46-62` by owner ion using the column pattern of
    `precalculate.py:277-294`; `lin_S`, `lin_A` are the per-type decay (+ absorption) blocks.
    Returns `(state indices, rhs rows, jv rows | None)`; validated against the full operators in
    tests/test_oracle.py."""
    ion_type = np.asarray(ion_type)
    index_S_i, index_A_j, _ = ground_states(ion_type, n_S_states, n_A_states)
    gs = np.where(ion_type != 0, index_A_j, index_S_i)
    procs = _active_pair_processes(energy_transfer)
    blocks = (np.asarray(lin_S, dtype=np.float64).reshape(n_S_states, n_S_states),
              np.asarray(lin_A, dtype=np.float64).reshape(n_A_states, n_A_states))
    idx, f_rows, jv_rows = [], [], []
    for q in ions:
        f, jv = _ion_rows(int(q), ion_type, gs, positions, cell, blocks, procs, y, v)
        idx.append(np.arange(gs[q], gs[q] + len(f)))
        f_rows.append(f)
        if jv is not None:
            jv_rows.append(jv)
    return (np.concatenate(idx), np.concatenate(f_rows),
            np.concatenate(jv_rows) if jv_rows else None)
