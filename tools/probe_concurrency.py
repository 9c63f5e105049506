"""Do solves of independent handles overlap on one device?  N handles (one lattice each), solved
one after the other and then from N host threads.  DEVELOPMENT TOOL."""
import sys
import threading
import time
sys.path.insert(0, '.')
import numpy as np                                          # noqa: E402
from simetuc_b200 import configs, lattice, precalculate    # noqa: E402
from simetuc_b200.engine import RateSystem                  # noqa: E402
import bench                                                # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cte = configs.yb_tm_cte(N_uc=15, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8, rtol=1e-8, atol=1e-15, N_steps=1000)
t_pulse, t_sol = bench.t_grids(cte)
systems = []
for k in range(n):
    lat = lattice.generate(cte, seed=3000 + k)
    sysk = RateSystem(lat, cte['energy_transfer'], device=0)      # one handle per lattice (not the cache)
    sysk.set_batch(1)
    sysk.set_decay_blocks(*precalculate.decay_blocks(cte))
    sysk.set_abs_blocks(*precalculate.absorption_blocks(cte))
    sysk.set_strengths([p.strength for p in cte['energy_transfer'].values() if len(p.type) == 2])
    systems.append(sysk)


def solve(sysk, out, k):
    _, _, pop = sysk.ground_states()
    t0 = time.perf_counter()
    yp, _, _ = sysk.integrate(t_pulse, pop.astype(float)[None, :], pulse=True, rtol=1e-8, atol=1e-15)
    ys, avg, st = sysk.integrate(t_sol, yp[:, -1], pulse=False, rtol=1e-8, atol=1e-15, want_y=False, want_avg=True)
    out[k] = (time.perf_counter() - t0, sysk.last_timing()[0])


out = [None] * n
solve(systems[0], out, 0)      # warm-up
t0 = time.perf_counter()
for k in range(n):
    solve(systems[k], out, k)
print('sequential: %.3f s total; per solve wall/kernel ms: %s' % (
    time.perf_counter() - t0, [(round(1e3 * a), round(b)) for a, b in out]))
t0 = time.perf_counter()
ths = [threading.Thread(target=solve, args=(systems[k], out, k)) for k in range(n)]
[t.start() for t in ths]
[t.join() for t in ths]
print('%d threads: %.3f s total; per solve wall/kernel ms: %s' % (
    n, time.perf_counter() - t0, [(round(1e3 * a), round(b)) for a, b in out]))


# ---- the same with the handle built (and closed) inside the thread, as the API path does -------
def build_and_solve(k, out):
    t0 = time.perf_counter()
    lat = lattice.generate(cte, seed=3100 + k)
    sysk = RateSystem(lat, cte['energy_transfer'], device=0)
    sysk.set_batch(1)
    sysk.set_decay_blocks(*precalculate.decay_blocks(cte))
    sysk.set_abs_blocks(*precalculate.absorption_blocks(cte))
    sysk.set_strengths([p.strength for p in cte['energy_transfer'].values() if len(p.type) == 2])
    t1 = time.perf_counter()
    if BARRIER is not None:
        BARRIER.wait()
    t1b = time.perf_counter()
    tmp = [None] * n
    solve(sysk, tmp, k)
    t2 = time.perf_counter()
    if BARRIER is not None:
        BARRIER.wait()
    t2b = time.perf_counter()
    sysk.close()
    out[k] = (round(1e3 * (t1 - t0)), round(1e3 * (t2 - t1b)), round(tmp[k][1]), round(1e3 * (time.perf_counter() - t2b)))


BARRIER = None
for label, par in (('sequential', False), ('threads', True), ('threads, phases separated by barriers', True)):
    BARRIER = threading.Barrier(n) if 'barriers' in label else None
    t0 = time.perf_counter()
    if par:
        ths = [threading.Thread(target=build_and_solve, args=(k, out)) for k in range(n)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    else:
        for k in range(n):
            build_and_solve(k, out)
    print('build+solve+close %s: %.3f s total; (build, solve wall, relax kernel, close) ms: %s' % (label, time.perf_counter() - t0, out))
