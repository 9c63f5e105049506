"""Golden vectors for the optimisation objective: the UNMODIFIED reference's
`DynamicsSolution._load_exp_data`, `_get_background`, `_interpolate_sim_data` and the error formula
of `_calc_errors` (`simetuc/simulations.py:354-527`) on the four decay curves shipped for the
config_file.cfg system (`expData/bNaYF4/decay_*_exc_NIR_980_25.0Yb_0.3Tm.txt`), against seeded
synthetic simulated curves.  TEST INFRASTRUCTURE; build container only (needs /root/reference).

    python -m oracle.make_golden_expdata      # writes tests/golden/expdata_objective.npz

Stored per curve: the raw two-column data of the file (the test writes it back to a temporary
expData tree), the reference's normalised data, its background offset, its interpolated simulated
curve and its error value.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

STATES = ['Tm_1D2', 'Tm_1G4', 'Tm_3H4', 'Yb_ES']


def synthetic_curve(t, k):
    rng = np.random.default_rng(9000 + k)
    tau_r, tau_d = 10.0 ** rng.uniform(-5, -4), 10.0 ** rng.uniform(-3.5, -2.5)
    return 1e-3 * (k + 1) * (np.exp(-t / tau_d) - np.exp(-t / tau_r)).clip(min=0) + 1e-9


def main():
    from oracle import ref_import
    ref_import.load()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import simetuc.simulations as ref_sim
    DS = ref_sim.DynamicsSolution
    cwd = os.getcwd()
    os.chdir(ref_import.REF_ROOT)          # the reference looks for ./simetuc/expData (simulations.py:362-365)
    out = {'states': np.array(STATES)}
    t_sol = np.logspace(-8, np.log10(0.18), 1000) - 1e-8
    out['t_sol'] = t_sol
    try:
        for k, state in enumerate(STATES):
            fname = 'decay_%s_exc_NIR_980_25.0Yb_0.3Tm.txt' % state
            path = os.path.join('simetuc', 'expData', 'bNaYF4', fname)
            raw = np.array([[float(x) for x in line.replace(',', ' ').replace(';', ' ').split()]
                            for line in open(path) if line.strip() and not line.startswith('#')
                            and len(line.replace(',', ' ').replace(';', ' ').split()) == 2])
            exp = DS._load_exp_data(fname, 'bNaYF4')
            assert exp is not None and exp.shape == raw.shape, (fname, None if exp is None else exp.shape, raw.shape)
            sim = synthetic_curve(t_sol, k)
            offset = DS._get_background(exp, sim)
            interp = DS._interpolate_sim_data(exp, sim, t_sol)
            err = 1 / np.mean(interp) * np.sqrt(1 / len(interp) * np.sum((interp - exp[:, 1] * np.max(interp)) ** 2))
            out['raw_' + state] = raw
            out['exp_' + state] = exp
            out['sim_' + state] = sim
            out['offset_' + state] = np.float64(offset)
            out['interp_' + state] = interp
            out['err_' + state] = np.float64(err)
            print(state, raw.shape, 'offset %.4g err %.6g' % (offset, err))
    finally:
        os.chdir(cwd)
    path = os.path.join(ROOT, 'tests', 'golden', 'expdata_objective.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
