"""BASELINE config 4 at scale: one large lattice (N_uc=40, 25 % Yb, 0.3 % Tm: 24 205 ions, 5.9e8
ordered pairs), pulse + 1000-point relaxation through the persistent grid integrator; reports
the algorithmic HBM traffic rate inside the solve.  DEVELOPMENT / MEASUREMENT TOOL."""
import json
import sys
import time
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs, lattice, precalculate   # noqa: E402
import bench                                               # noqa: E402

N_uc = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
cte = configs.yb_tm_cte(N_uc=N_uc, S_conc=25.0, A_conc=0.3, power_dens=1e7, t_pulse=1e-8,
                        rtol=rtol, atol=1e-15, N_steps=1000)
t0 = time.time()
lat = lattice.generate(cte, seed=4000)
t1 = time.time()
system = precalculate.build_system(cte, lat, n_sys=1)
system.set_abs_blocks(*precalculate.absorption_blocks(cte), system=0)
t2 = time.time()
_, _, pop = system.ground_states()
t_pulse, t_sol = bench.t_grids(cte)
y0 = pop.astype(float)[None, :]
out = {'N_uc': N_uc, 'ions': int(system.n_ions), 'states': int(system.N), 'pairs': int(system.n_pairs),
       'stream_GB': system.stream_bytes / 1e9, 'lattice_s': t1 - t0, 'build_s': t2 - t1, 'rtol': rtol}
tot_ms, tot_bytes = 0.0, 0.0
for name, t, pulse, y in (('pulse', t_pulse, True, y0), ('relax', t_sol, False, None)):
    if y is None:
        y = yp[:, -1]
    ys, avg, st = system.integrate(t, y, pulse=pulse, rtol=rtol, atol=1e-15, want_y=(name == 'pulse'),
                                   want_avg=True, nsteps=100000)
    if name == 'pulse':
        yp = ys
    ms, _ = system.last_timing()
    s = st[0]
    b = s['n_rhs'] * (system.bytes_per_pair * system.n_pairs + 16 * system.N) + \
        s['n_jv'] * (system.bytes_per_pair * system.n_pairs + 24 * system.N)
    out[name] = {'ms': ms, 'steps': s['n_steps'], 'n_rhs': s['n_rhs'], 'n_jv': s['n_jv'], 'status': s['status'],
                 'algorithmic_GBps': b / (ms * 1e-3) / 1e9, 'ms_per_pass': ms / max(1, s['n_rhs'] + s['n_jv'])}
    tot_ms += ms
    tot_bytes += b
out['total_s'] = tot_ms / 1e3
out['algorithmic_GBps'] = tot_bytes / (tot_ms * 1e-3) / 1e9
out['final_avg'] = [float(v) for v in avg[0, -1]]
print(json.dumps(out))
