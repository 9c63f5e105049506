"""Where the C3 sweep (120 lattices, 16 in flight) spends its wall time (DEVELOPMENT TOOL)."""
import sys
import threading
import time
sys.path.insert(0, '.')
from simetuc_b200 import configs, engine, lattice, precalculate   # noqa: E402
from simetuc_b200.simulations import Simulations                  # noqa: E402

acc, lock = {}, threading.Lock()


def wrap(obj, name, label=None):
    orig = getattr(obj, name)
    label = label or name

    def timed(*a, **k):
        t0 = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            with lock:
                d = acc.setdefault(label, [0.0, 0])
                d[0] += time.perf_counter() - t0
                d[1] += 1
    setattr(obj, name, timed)


wrap(lattice, 'generate')
wrap(engine.RateSystem, '__init__', 'RateSystem.__init__')
wrap(engine.RateSystem, 'integrate_dynamics')
wrap(precalculate, 'acquire_system')
concs, n_uc = [], []
for _seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    for S, nuc in ((10.0, 20), (25.0, 15), (50.0, 12)):
        for A in (0.1, 0.3, 0.5, 1.0, 2.0):
            concs.append((S, A))
            n_uc.append(nuc)
cte = configs.yb_tm_cte(N_uc=15, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8, rtol=1e-8, atol=1e-15,
                        N_steps=1000)
cte['lattice_seed'] = 3000
for streams in (16, 32):
    sim = Simulations(cte, devices=[0], streams_per_device=streams)
    sim.simulate_concentration_dependence(concs[:2], n_uc[:2], dynamics=True, keep_y=False)
    acc.clear()
    precalculate._LATTICE_CACHE.clear()
    t0 = time.perf_counter()
    sol = sim.simulate_concentration_dependence(concs, n_uc, dynamics=True, keep_y=False)
    dt = time.perf_counter() - t0
    print('streams %d: %d lattices in %.2f s = %.1f solves/s | thread-seconds:' % (streams, len(concs), dt, len(concs) / dt),
          ' '.join('%s=%.2f(%d)' % (k, v[0], v[1]) for k, v in acc.items()), flush=True)
    precalculate.clear_system_pool()
