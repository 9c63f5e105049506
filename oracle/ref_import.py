"""Import the UNMODIFIED reference (`/root/reference/simetuc`) in this container.

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists only in the build container, never on
the GPU box, so nothing in the `-m gpu` tests, `smoke()` or `bench.py` may import this
module.  It is used by `oracle/make_golden.py` (to generate the committed fixtures under
`tests/golden/`) and by the container-only tests that pin `oracle/oracle_np.py` against
the live reference.

The reference needs packages that are not installed here (h5py, ruamel.yaml, ase,
settings_parser, matplotlib); its *arithmetic* for the hot path (`odesolver.py:33-80`,
`precalculate.py:51-415,603-623`, `lattice.py:259-374`) does not use them, so empty stub
modules are registered in `sys.modules` before the import.
"""
import os
import sys
import types
import warnings

REF_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'simetuc'))


class _Tolerant:
    """Object that accepts any constructor args / attribute access (schema stubs)."""
    optional = exclusive = None

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Tolerant()

    def __call__(self, *a, **k):
        return _Tolerant()

    def __iter__(self):
        return iter(())


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def install_stubs():
    _stub('h5py')
    ru = _stub('ruamel')
    ru.yaml = _stub('ruamel.yaml')
    ase = _stub('ase', Atoms=object)
    ase.io = _stub('ase.io')
    ase.spacegroup = _stub('ase.spacegroup', crystal=None)

    class Settings(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    _stub('settings_parser', Settings=Settings,
          SettingsValueError=type('SettingsValueError', (Exception,), {}),
          SettingsExtraValueWarning=type('SettingsExtraValueWarning', (UserWarning,), {}),
          Value=_Tolerant, DictValue=_Tolerant, Kind=_Tolerant)
    mpl = _stub('matplotlib', colors=_Tolerant(), figure=_Tolerant())
    mpl.pyplot = _stub('matplotlib.pyplot')
    mt = _stub('mpl_toolkits')
    mt.mplot3d = _stub('mpl_toolkits.mplot3d', Axes3D=object)


_cache = {}


def load():
    """Return a namespace with the reference modules `odesolver`, `precalculate`,
    `lattice`, `util` imported from /root/reference (stubs installed first)."""
    if _cache:
        return _cache['ns']
    if not available():
        raise RuntimeError('reference tree not present (only exists in the build container)')
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import simetuc.util as util
        import simetuc.odesolver as odesolver
        import simetuc.precalculate as precalculate
        import simetuc.lattice as lattice
    ns = types.SimpleNamespace(util=util, odesolver=odesolver,
                               precalculate=precalculate, lattice=lattice)
    _cache['ns'] = ns
    return ns
