"""G = 1 vs 4 of the cluster integrator on the 240S/108A pulse (DEVELOPMENT TOOL)."""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
from simetuc_b200 import configs                      # noqa: E402
from tests.helpers import load_golden                 # noqa: E402
from tests.test_gpu_parity import make_system         # noqa: E402
os.environ['STB_PCG'] = os.environ.get('STB_PCG', '4')
os.environ['STB_LIN_RMS'] = '2e-3'
system = make_system(load_golden('240S_108A'), configs.test_cte())
_, _, pop = system.ground_states()
t = np.array([0.0, 1e-8])
res = {}
for G in (1, 2, 4, 8):
    os.environ['STB_GROUP_CTAS'] = str(G)
    y, _, st = system.integrate(t, pop.astype(float), pulse=True, rtol=1e-8, atol=1e-15)
    prof = system.last_profile()
    print(G, 'G used', prof['G'], 'steps', st[0]['n_steps'], 'jv', st[0]['n_jv'], 'lin', st[0]['n_lin'], 'rej', st[0]['n_rejected'],
          'pc_group', system.get_option('pc_group'))
    if G in res:
        print('  repeat identical:', np.array_equal(res[G], y[0, -1]))
    res[G] = y[0, -1]
for G in (2, 4, 8):
    d = np.abs(res[G] - res[1]) / (1e-9 * np.abs(res[1]) + 1e-18)
    k = int(np.argmax(d))
    print('G', G, 'max ratio', d.max(), 'at', k, res[1][k], res[G][k])
