"""Numerical prototype (numpy) of the device integrator: variable-order BDF (NDF, orders 1-5)
+ Newton-Krylov (left-preconditioned GMRES, per-ion block-Jacobi).  DEVELOPMENT TOOL: it models
the algorithm implemented in `simetuc_b200/csrc/` so that step/iteration counts and accuracy can
be studied on the CPU; it is not imported by the package and is not part of the product path.
"""
import numpy as np

MAX_ORDER = 5
NEWTON_MAXITER = 4
MIN_FACTOR = 0.2
MAX_FACTOR = 10.0
KAPPA = np.array([0, -0.1850, -1 / 9, -0.0823, -0.0415, 0])
GAMMA = np.hstack((0, np.cumsum(1 / np.arange(1, MAX_ORDER + 1))))
ALPHA = (1 - KAPPA) * GAMMA
ERR_CONST = KAPPA * GAMMA + 1 / np.arange(1, MAX_ORDER + 2)


def compute_R(order, factor):
    I = np.arange(1, order + 1)[:, None]
    J = np.arange(1, order + 1)
    M = np.zeros((order + 1, order + 1))
    M[1:, 1:] = (I - 1 - factor * J) / I
    M[0] = 1
    return np.cumprod(M, axis=0)


def change_D(D, order, factor):
    R = compute_R(order, factor)
    U = compute_R(order, 1)
    RU = R.dot(U)
    D[:order + 1] = np.dot(RU.T, D[:order + 1])


def rms(x):
    return np.linalg.norm(x) / np.sqrt(x.size)


class Stats:
    def __init__(self):
        self.nfev = self.njv = self.nsteps = self.nrej = self.nnewt_fail = self.npsetup = 0
        self.nlin = 0
        self.krylov_hist = []


def gmres(A, b, Minv, tol, maxit, stats):
    """Left-preconditioned GMRES on M^-1 A x = M^-1 b, x0 = 0; stops when the preconditioned
    residual 2-norm <= tol.  Returns x, n_iter, converged."""
    n = b.size
    r = Minv(b)
    beta = np.linalg.norm(r)
    if beta <= tol:
        return np.zeros(n), 0, True
    V = np.zeros((maxit + 1, n))
    H = np.zeros((maxit + 1, maxit))
    V[0] = r / beta
    g = np.zeros(maxit + 1)
    g[0] = beta
    cs = np.zeros(maxit)
    sn = np.zeros(maxit)
    k = 0
    conv = False
    for j in range(maxit):
        w = Minv(A(V[j]))
        stats.njv += 1
        # classical Gram-Schmidt, twice (what the device does: batched dots)
        for _ in range(2):
            h = V[:j + 1] @ w
            w = w - h @ V[:j + 1]
            H[:j + 1, j] += h
        H[j + 1, j] = np.linalg.norm(w)
        if H[j + 1, j] > 0:
            V[j + 1] = w / H[j + 1, j]
        for i in range(j):
            t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
            H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
            H[i, j] = t
        d = np.hypot(H[j, j], H[j + 1, j])
        cs[j], sn[j] = H[j, j] / d, H[j + 1, j] / d
        H[j, j] = d
        H[j + 1, j] = 0
        g[j + 1] = -sn[j] * g[j]
        g[j] = cs[j] * g[j]
        k = j + 1
        if abs(g[j + 1]) <= tol:
            conv = True
            break
    yk = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
    return yk @ V[:k], k, conv


def solve(fun, jv_factory, precond_factory, t_arr, y0, rtol, atol, first_step=None,
          max_steps=200000, krylov_max=30, eps_lin=0.05, verbose=False, carry_rate=False,
          verify_every=0, guess=False):
    """Integrate y' = fun(y) and return y at t_arr (dense output by BDF interpolation).
    jv_factory(y) -> (v -> J(y) v); precond_factory(y, c) -> (r -> M^-1 r), M ~ I - c J."""
    st = Stats()
    t_arr = np.asarray(t_arr, dtype=float)
    n = y0.size
    out = np.zeros((len(t_arr), n))
    out[0] = y0
    t = t_arr[0]
    t_end = t_arr[-1]
    y = y0.astype(float).copy()
    f0 = fun(y)
    st.nfev += 1
    if first_step is None:
        scale = atol + np.abs(y) * rtol
        d0 = rms(y / scale)
        d1 = rms(f0 / scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        y1 = y + h0 * f0
        f1 = fun(y1)
        st.nfev += 1
        d2 = rms((f1 - f0) / scale) / h0
        h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** 0.5
        h = min(100 * h0, h1, t_end - t)
    else:
        h = first_step
    D = np.zeros((MAX_ORDER + 3, n))
    D[0] = y
    D[1] = f0 * h
    order = 1
    n_equal = 0
    newton_tol = max(10 * np.finfo(float).eps / rtol, min(0.03, rtol ** 0.5))
    k_out = 1
    gfac = 0.0
    gfac2 = False
    n_same = 0
    crate = 1.0              # Newton convergence rate carried across steps (LSODA / CVODE practice)
    since_verify = 0
    st.krylov_by_it = [[] for _ in range(NEWTON_MAXITER)]
    st.rates = []
    while k_out < len(t_arr) and st.nsteps < max_steps:
        # ---- one accepted step
        step_accepted = False
        while not step_accepted:
            t_new = t + h
            if t_new > t_end * (1 + 1e-14) and (t_new - t_end) > 1e-14 * abs(t_end):
                t_new = t_end
                change_D(D, order, (t_new - t) / h)
                gfac *= ((t_new - t) / h) ** (order + 1)
                n_equal = 0
            h = t_new - t
            y_pred = D[:order + 1].sum(0)
            scale = atol + rtol * np.abs(y_pred)
            psi = (D[1:order + 1].T @ GAMMA[1:order + 1]) / ALPHA[order]
            c = h / ALPHA[order]
            # Newton-Krylov
            d = (gfac * (D[order + 1] + (D[order + 2] if (guess == 2 and gfac2) else 0.0))) if (guess and gfac != 0.0) else np.zeros(n)
            yk = y_pred + d
            converged = False
            dy_norm_old = None
            Minv = None
            for it in range(NEWTON_MAXITER):
                f = fun(yk)
                st.nfev += 1
                if not np.all(np.isfinite(f)):
                    break
                if Minv is None:
                    P = precond_factory(yk, c)
                    st.npsetup += 1
                    Jv = jv_factory(yk)
                    # scaled system: unknown x = dy/scale; operator S^-1 (I - cJ) S
                    A = lambda x, Jv=Jv: x - c * Jv(x * scale) / scale
                    Minv = lambda r, P=P: P(r * scale) / scale
                res = (c * f - psi - d) / scale
                tol_lin = eps_lin * newton_tol * np.sqrt(n)
                x, kk, ok = gmres(A, res, Minv, tol_lin, krylov_max, st)
                st.nlin += 1
                st.krylov_hist.append(kk)
                st.krylov_by_it[it].append(kk)
                dy = x * scale
                dy_norm = rms(x)
                rate = None if dy_norm_old is None else dy_norm / dy_norm_old
                if rate is not None:
                    crate = max(0.2 * crate, rate)
                    st.rates.append(rate)
                    since_verify = 0
                if rate is not None and (rate >= 1 or
                                         rate ** (NEWTON_MAXITER - it) / (1 - rate) * dy_norm > newton_tol):
                    break
                yk = yk + dy
                d = d + dy
                if dy_norm == 0 or (rate is not None and rate / (1 - rate) * dy_norm < newton_tol) \
                        or (rate is None and dy_norm < 0.1 * newton_tol):
                    converged = True
                    break
                if carry_rate and rate is None and (verify_every == 0 or since_verify < verify_every) \
                        and 1.5 * min(1.0, crate) * dy_norm < newton_tol:
                    converged = True
                    since_verify += 1
                    break
                dy_norm_old = dy_norm
            if not converged:
                st.nnewt_fail += 1
                h *= 0.5
                change_D(D, order, 0.5)
                gfac *= 0.5 ** (order + 1)
                n_equal = 0
                continue
            n_iter = it + 1
            safety = 0.9 * (2 * NEWTON_MAXITER + 1) / (2 * NEWTON_MAXITER + n_iter)
            scale = atol + rtol * np.abs(yk)
            err = ERR_CONST[order] * d
            err_norm = rms(err / scale)
            if err_norm > 1:
                st.nrej += 1
                factor = max(MIN_FACTOR, safety * err_norm ** (-1 / (order + 1)))
                h *= factor
                change_D(D, order, factor)
                gfac *= factor ** (order + 1)
                n_equal = 0
            else:
                step_accepted = True
        st.nsteps += 1
        n_equal += 1
        gfac2 = (gfac == 1.0)
        gfac = 1.0
        t_old = t
        t = t_new
        y = yk
        D[order + 2] = d - D[order + 1]
        D[order + 1] = d
        for i in reversed(range(order + 1)):
            D[i] += D[i + 1]
        # dense output for every requested time in (t_old, t]
        while k_out < len(t_arr) and t_arr[k_out] <= t * (1 + 1e-15):
            tq = t_arr[k_out]
            if tq >= t:
                out[k_out] = D[0]
            else:
                shifts = t - h * np.arange(order)
                denom = h * (1 + np.arange(order))
                p = np.cumprod((tq - shifts) / denom)
                out[k_out] = D[0] + p @ D[1:order + 1]
            k_out += 1
        if n_equal < order + 1:
            continue
        err_m = rms(ERR_CONST[order - 1] * D[order] / scale) if order > 1 else np.inf
        err_p = rms(ERR_CONST[order + 1] * D[order + 2] / scale) if order < MAX_ORDER else np.inf
        norms = np.array([err_m, err_norm, err_p])
        with np.errstate(divide='ignore'):
            factors = norms ** (-1 / np.arange(order, order + 3))
        delta = int(np.argmax(factors)) - 1
        order += delta
        factor = min(MAX_FACTOR, safety * np.max(factors))
        h *= factor
        change_D(D, order, factor)
        gfac = factor ** (order + 1) if delta == 0 else 0.0
        n_equal = 0
        if verbose and st.nsteps % 200 == 0:
            print('step', st.nsteps, 't', t, 'h', h, 'order', order)
    return out, st
