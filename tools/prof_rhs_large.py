"""RHS + Jv kernels on the HBM-bound lattice (N_uc=40, 25% Yb, 0.3% Tm) - ncu target."""
import ctypes as C
import sys
sys.path.insert(0, '.')
import torch
from simetuc_b200 import _lib, configs, lattice, precalculate
cte = configs.yb_tm_cte(N_uc=40, S_conc=25.0, A_conc=0.3)
lat = lattice.generate(cte, seed=4000)
system = precalculate.build_system(cte, lat, n_sys=1)
N = system.N
y = torch.rand(1, N, dtype=torch.float64, device='cuda') * 1e-3
v = torch.rand(1, N, dtype=torch.float64, device='cuda')
out = torch.empty_like(y)
ms = C.c_float()
for _ in range(2):
    _lib.check(system.lib.stb_rhs_dev(system.handle, 0, 1, 1, y.data_ptr(), out.data_ptr(), 3, C.byref(ms)))
    r = ms.value
    _lib.check(system.lib.stb_jv_dev(system.handle, 0, 1, 1, y.data_ptr(), v.data_ptr(), out.data_ptr(), 3, C.byref(ms)))
print('pairs', system.n_pairs, 'N', N, 'rhs ms', r, 'jv ms', ms.value)
