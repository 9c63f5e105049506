"""Where the end-to-end sweep call spends its time: every C-ABI call timed on the host next to the
kernel time the library reports (DEVELOPMENT TOOL)."""
import sys
import time
sys.path.insert(0, '.')
import torch                                      # noqa: E402,F401
import bench                                      # noqa: E402
from simetuc_b200 import _lib, engine, precalculate             # noqa: E402
from simetuc_b200.simulations import Simulations  # noqa: E402

acc = {}


def wrap(cls, name):
    orig = getattr(cls, name)

    def timed(*a, **k):
        t0 = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            d = acc.setdefault(name, [0.0, 0])
            d[0] += time.perf_counter() - t0
            d[1] += 1
    setattr(cls, name, timed)


for nm in ('__init__', 'set_batch', 'set_decay_blocks', 'set_abs_blocks', 'set_strengths', 'ground_states',
           'integrate', 'integrate_dynamics', 'close'):
    if hasattr(engine.RateSystem, nm):
        wrap(engine.RateSystem, nm)
for nm in ('decay_blocks', 'absorption_blocks', 'get_lattice', 'build_system', 'acquire_system'):
    if hasattr(precalculate, nm):
        wrap(precalculate, nm)
powers = bench.powers_list()
sim = Simulations(bench.make_cte(), devices=[0])
sim.cte['lattice_seed'] = 2000
for rep in range(8):
    acc.clear()
    t0 = time.perf_counter()
    sol = sim.simulate_power_dependence(powers, keep_y=False)
    _ = sol.populations
    dt = 1e3 * (time.perf_counter() - t0)
    print('e2e call: %.1f ms -> %.0f solves/s |' % (dt, 32e3 / dt),
          ' '.join('%s=%.1f(%d)' % (k, 1e3 * v[0], v[1]) for k, v in acc.items()), flush=True)
