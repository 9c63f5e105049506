"""Krylov iterations per linear solve of the numpy model (tools/proto_bdf.py) for block preconditioners over
groups of sensitizers: greedy coupling-driven groups (S4 .. S32), k-d boxes (kd4 .. kd32), two-level variants.
DEVELOPMENT TOOL: python tools/proto_groups.py <variant> [S_conc]  (results: profiles/r02_groups.txt)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from oracle import oracle_np as onp
from simetuc_b200 import configs, lattice
from tests.helpers import oracle_operands
from tools.proto_run import make_problem, block_slices
from tools import proto_bdf as pb
variant = sys.argv[1]
S_conc = float(sys.argv[2]) if len(sys.argv) > 2 else 25.0
cte = configs.yb_tm_cte(N_uc=10, S_conc=S_conc, A_conc=0.3, power_dens=1e7)
lat = lattice.generate_arrays(cte, seed=1012)
g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'], 'info': np.array([2, 9])}
op = oracle_operands(g, cte)
tf = onp.t_simulation(cte['decay'])
t_pulse, t_sol = onp.time_grids(1e-8, tf, 1000)
y0 = op['initial_population'].astype(float)
N = op['N']
sl = block_slices(op)
ion_type = lat['ion_type']; dist = lat['dist_array']
n_ions = len(ion_type)
def groups_S(maxsz):
    d = dist.copy(); np.fill_diagonal(d, np.inf)
    w = d ** -6.0
    pairs = sorted(((w[i, j], i, j) for i in range(n_ions) for j in range(i + 1, n_ions) if ion_type[i] == 0 and ion_type[j] == 0), reverse=True)
    grp = list(range(n_ions)); members = {i: [i] for i in range(n_ions)}
    for wij, i, j in pairs:
        gi, gj = grp[i], grp[j]
        if gi == gj: continue
        if len(members[gi]) + len(members[gj]) > maxsz: continue
        for m in members[gj]: grp[m] = gi
        members[gi] += members.pop(gj)
    return list(members.values())
def make_pf(grps, jvf, coarse=None):
    idx_groups = [np.concatenate([np.arange(sl[q][0], sl[q][0] + sl[q][1]) for q in grp]) for grp in grps]
    Pm = None
    if coarse is not None:
        cols = []
        for grp in coarse:
            v = np.zeros(N)
            for q in grp:
                if ion_type[q] == 0:
                    v[sl[q][0]] = -1.0; v[sl[q][0] + 1] = 1.0
            if np.any(v): cols.append(v)
        Pm = np.column_stack(cols)
    def pf(y, c):
        Jv = jvf(y)
        J = np.column_stack([Jv(e) for e in np.eye(N)])
        A = np.eye(N) - c * J
        invs = [np.linalg.inv(A[np.ix_(ix, ix)]) for ix in idx_groups]
        if Pm is not None:
            Ac_inv = np.linalg.inv(Pm.T @ A @ Pm)
        def apply(r):
            out = np.empty_like(r)
            for ix, Bi in zip(idx_groups, invs): out[ix] = Bi @ r[ix]
            if Pm is not None:
                # multiplicative two-level: coarse correction of the remaining residual
                out = out + Pm @ (Ac_inv @ (Pm.T @ (r - A @ out)))
            return out
        return apply
    return pf
fun, jvf, pf0 = make_problem(op, True)
yp, st1 = pb.solve(fun, jvf, pf0, t_pulse, y0, 1e-8, 1e-15, eps_lin=20, carry_rate=True, verify_every=16, guess=True)
fun, jvf, pf0 = make_problem(op, False)
pos = np.asarray(lat['doped_lattice'], dtype=float)
def kd_groups(leaf):
    S_ids = [q for q in range(n_ions) if ion_type[q] == 0]
    out = []
    def rec(ids):
        m = len(ids)
        if m <= leaf:
            out.append(list(ids)); return
        p = pos[ids]
        ax = int(np.argmax(p.max(0) - p.min(0)))
        order = np.argsort(p[:, ax], kind='stable')
        n_leaves = (m + leaf - 1) // leaf
        left = leaf * ((n_leaves + 1) // 2)
        ids = np.asarray(ids)[order]
        rec(list(ids[:left])); rec(list(ids[left:]))
    rec(S_ids)
    return out + [[q] for q in range(n_ions) if ion_type[q] != 0]
single = [[q] for q in range(n_ions)]
S_all = [q for q in range(n_ions) if ion_type[q] == 0]
variants = {
 'single': (single, None),
 'Squad': (groups_S(4), None),
 'S8': (groups_S(8), None),
 'S16': (groups_S(16), None),
 'Sall': ([S_all] + [[q] for q in range(n_ions) if ion_type[q] != 0], None),
 'Squad+coarse4': (groups_S(4), groups_S(4)),
 'Squad+coarse8': (groups_S(4), groups_S(8)),
 'single+coarse4': (single, groups_S(4)),
 'Squad+global': (groups_S(4), [S_all]),
 'kd4': (kd_groups(4), None), 'kd8': (kd_groups(8), None), 'kd16': (kd_groups(16), None), 'kd32': (kd_groups(32), None), 'S32': (groups_S(32), None),
}
grps, coarse = variants[variant]
t0 = time.time()
ys, st2 = pb.solve(fun, jvf, make_pf(grps, jvf, coarse), t_sol, yp[-1], 1e-8, 1e-15, eps_lin=20, carry_rate=True, verify_every=16, guess=True)
kh = np.array(st2.krylov_hist)
q = len(kh) // 4
print('%-16s S %.0f%% ions %d N %d: steps %d nfev %d njv %d  its/solve %.2f  by quarter %s (%.0fs)' % (variant, S_conc, n_ions, N, st2.nsteps, st2.nfev, st2.njv, st2.njv / st2.nlin,
      [round(float(kh[i*q:(i+1)*q].mean()), 2) for i in range(4)], time.time() - t0), flush=True)
