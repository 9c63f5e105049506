"""Drive tools/proto_bdf.py on oracle operands and compare with the scipy-LSODA oracle.
DEVELOPMENT TOOL (not part of the product path)."""
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, '.')
from oracle import oracle_np as onp            # noqa: E402
from simetuc_b200 import configs, lattice      # noqa: E402
from tests.helpers import load_golden, oracle_operands   # noqa: E402
from tools import proto_bdf                    # noqa: E402


def block_slices(op):
    sl = []
    for s, a in zip(op['index_S_i'], op['index_A_j']):
        if s != -1:
            sl.append((int(s), op['sS']))
        elif a != -1:
            sl.append((int(a), op['sA']))
    return sl


def make_problem(op, pulse):
    lin = (op['abs'] + op['decay']).tocsr() if pulse else op['decay'].tocsr()
    ET = op['ET'].tocsr()
    ni = op['N_indices'][:, 0].astype(np.int64)
    nj = op['N_indices'][:, 1].astype(np.int64)
    N = op['N']
    slices = block_slices(op)
    # structure for the intra-ion block Jacobian: for interaction r, ET column entries
    ETc = ET.tocsc()
    rows4 = ETc.indices.reshape(-1, 4) if ET.shape[1] else np.zeros((0, 4), np.int64)
    # scipy sorts indices in each column; recover w and positions robustly from data
    data4 = ETc.data.reshape(-1, 4) if ET.shape[1] else np.zeros((0, 4))
    ion_of_state = np.zeros(N, dtype=np.int64)
    for q, (gs, n) in enumerate(slices):
        ion_of_state[gs:gs + n] = q

    def fun(y):
        return lin.dot(y) + ET.dot(y[ni] * y[nj])

    def jv_factory(y):
        yi, yj = y[ni], y[nj]
        return lambda v: lin.dot(v) + ET.dot(yj * v[ni] + yi * v[nj])

    lin_d = lin.toarray() if N <= 3000 else None

    def precond_factory(y, c):
        # block-diagonal of I - c*J(y)
        M = ET.shape[1]
        rr, cc, vv = [], [], []
        if M:
            for col_idx, other in ((ni, nj), (nj, ni)):
                ion_c = ion_of_state[col_idx]
                for k in range(4):
                    same = ion_of_state[rows4[:, k]] == ion_c
                    rr.append(rows4[same, k])
                    cc.append(col_idx[same])
                    vv.append(data4[same, k] * y[other][same])
            Jintra = sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))),
                                   shape=(N, N))
        else:
            Jintra = sp.csr_matrix((N, N))
        Jb = (lin + Jintra).tocsr()
        inv_blocks = []
        for gs, n in slices:
            B = np.eye(n) - c * Jb[gs:gs + n, gs:gs + n].toarray()
            inv_blocks.append(np.linalg.inv(B))

        def apply(r):
            out = np.empty_like(r)
            for (gs, n), Bi in zip(slices, inv_blocks):
                out[gs:gs + n] = Bi @ r[gs:gs + n]
            return out
        return apply

    return fun, jv_factory, precond_factory


def compare(y, yref, label):
    err = np.abs(y - yref) / (1e-4 * np.abs(yref) + 1e-12)
    print('%s: max parity metric (<=1 passes) %.3g ; max rel err where |y|>1e-9: %.3g' % (
        label, err.max(), (np.abs(y - yref) / np.abs(yref))[np.abs(yref) > 1e-9].max()))


def run(op, cte, t_pulse_w, rtol=1e-8, atol=1e-15, n_steps=1000, do_ref=True, krylov_max=30, ref=None, **kw_solve):
    tf = onp.t_simulation(cte['decay'])
    t_pulse, t_sol = onp.time_grids(t_pulse_w, tf, n_steps)
    y0 = op['initial_population'].astype(float)
    res = {}
    t0 = time.time()
    fun, jvf, pf = make_problem(op, True)
    yp, st1 = proto_bdf.solve(fun, jvf, pf, t_pulse, y0, rtol, atol, krylov_max=krylov_max, **kw_solve)
    fun, jvf, pf = make_problem(op, False)
    ys, st2 = proto_bdf.solve(fun, jvf, pf, t_sol, yp[-1], rtol, atol, krylov_max=krylov_max, **kw_solve)
    print('proto: %.1fs' % (time.time() - t0))
    for name, st in (('pulse', st1), ('relax', st2)):
        kh = np.array(st.krylov_hist)
        print(' %s: steps %d rej %d newton_fail %d nfev %d njv %d lin solves %d  krylov/solve mean %.1f max %d' % (
            name, st.nsteps, st.nrej, st.nnewt_fail, st.nfev, st.njv, st.nlin,
            kh.mean() if len(kh) else 0, kh.max() if len(kh) else 0))
        print('   krylov by newton it:', [(len(k), round(float(np.mean(k)), 2)) for k in st.krylov_by_it if k],
              ' rates: n %d max %.2g median %.2g' % (len(st.rates), max(st.rates, default=0), np.median(st.rates) if st.rates else 0))
    if ref is not None:
        compare(yp[-1], ref[0], 'pulse end')
        compare(ys, ref[1], 'relax')
        res['ref'] = ref
    elif do_ref:
        t0 = time.time()
        kw = dict(rtol=rtol, atol=atol, nsteps=100000)
        yp_ref = onp.solve_pulse(t_pulse, y0, op['abs'], op['decay'], op['ET'], op['N_indices'],
                                 op['jac_indices'], **kw)
        ys_ref = onp.solve_relax(t_sol, yp_ref[-1], op['decay'], op['ET'], op['N_indices'],
                                 op['jac_indices'], **kw)
        print('lsoda: %.1fs' % (time.time() - t0))
        compare(yp[-1], yp_ref[-1], 'pulse end')
        compare(ys, ys_ref, 'relax')
        res['ref'] = (yp_ref[-1], ys_ref)
    res['y'] = ys
    return res


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'fixture'
    if which == 'fixture':
        g = load_golden('240S_108A')
        cte = configs.sim_cte()
        op = oracle_operands(g, cte)
        run(op, cte, 1e-8)
    elif which == 'tiny':
        g = load_golden('sim_2S_2A')
        cte = configs.sim_cte()
        op = oracle_operands(g, cte)
        run(op, cte, 1e-8)
    elif which == 'ybtm':
        N_uc = int(sys.argv[2]) if len(sys.argv) > 2 else 15
        P = float(sys.argv[3]) if len(sys.argv) > 3 else 1e7
        S = float(sys.argv[4]) if len(sys.argv) > 4 else 5.0
        A = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
        cte = configs.yb_tm_cte(N_uc=N_uc, S_conc=S, A_conc=A, power_dens=P)
        lat = lattice.generate_arrays(cte, seed=2001)
        g = {'ion_type': lat['ion_type'], 'dist_array': lat['dist_array'],
             'info': np.array([2, 9])}
        op = oracle_operands(g, cte)
        print('N =', op['N'], 'M =', op['N_indices'].shape[0])
        run(op, cte, 1e-8)
