"""RHS + Jv kernels on the HBM-bound lattice (N_uc=40, 25% Yb, 0.3% Tm) - timing and ncu target.
Usage: prof_rhs_large.py [modes]   modes = comma list of STB_TMA values (default "0,1"), each
optionally followed by :kc=K :st=S :diag=D :grid=G (ring geometry / diagnostics of stb_tma.cuh)."""
import ctypes as C
import os
import sys
sys.path.insert(0, '.')
import torch
from simetuc_b200 import _lib, configs, lattice, precalculate
modes = (sys.argv[1] if len(sys.argv) > 1 else '0,1').split(',')
n_uc = int(os.environ.get('PROF_NUC', '40'))
cte = configs.yb_tm_cte(N_uc=n_uc, S_conc=25.0, A_conc=0.3)
lat = lattice.generate(cte, seed=4000)
system = precalculate.build_system(cte, lat, n_sys=1)
N = system.N
y = torch.rand(1, N, dtype=torch.float64, device='cuda') * 1e-3
v = torch.rand(1, N, dtype=torch.float64, device='cuda')
out = torch.empty_like(y)
ref = None
ms = C.c_float()
for spec in modes:
    mode, *opts = spec.split(':')
    os.environ['STB_TMA'] = mode
    for var in ('STB_TMA_KC', 'STB_TMA_STAGES', 'STB_TMA_DIAG', 'STB_TMA_GRID'):
        os.environ.pop(var, None)
    for o in opts:
        k, val = o.split('=')
        os.environ[{'kc': 'STB_TMA_KC', 'st': 'STB_TMA_STAGES', 'diag': 'STB_TMA_DIAG', 'grid': 'STB_TMA_GRID'}[k]] = val
    for _ in range(2):
        _lib.check(system.lib.stb_rhs_dev(system.handle, 0, 1, 1, y.data_ptr(), out.data_ptr(), 5, C.byref(ms)))
        r = ms.value
        f = out.clone()
        _lib.check(system.lib.stb_jv_dev(system.handle, 0, 1, 1, y.data_ptr(), v.data_ptr(), out.data_ptr(), 5, C.byref(ms)))
    gb = system.n_pairs * 8 / 1e9
    if ref is None:
        ref = (f, out.clone())
    err = max(float((f - ref[0]).abs().max() / ref[0].abs().max()), float((out - ref[1]).abs().max() / ref[1].abs().max()))
    print('STB_TMA=%-22s pairs %d N %d rhs %.4f ms %.0f GB/s  jv %.4f ms %.0f GB/s  max rel diff vs first mode %.2e' % (
        spec, system.n_pairs, N, r, gb / r * 1e3, ms.value, gb / ms.value * 1e3, err))
if os.environ.get('PROF_READ_REF'):
    # read-only streaming reference: sum of a 4.3 GB f64 tensor (torch reduction), best of 5
    big = torch.rand(1 << 29, dtype=torch.float64, device='cuda')
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); big.sum(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('torch.sum over %.2f GB: %.4f ms = %.0f GB/s (read-only reference)' % (big.numel() * 8 / 1e9, best, big.numel() * 8 / best / 1e6))
