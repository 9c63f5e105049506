"""ctypes binding of the C ABI (`include/simetuc_b200.h`, built from `csrc/` into
`libsimetuc_b200.so`).  There is NO CPU fallback: if the library is missing or no CUDA device is
present, every compute entry point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('STB_LIB') or os.path.join(_HERE, 'libsimetuc_b200.so')

STB_MAX_STATES = 12
STB_MAX_PROC = 32


class StbProcess(C.Structure):
    _fields_ = [('donor_type', C.c_int32), ('partner_type', C.c_int32),
                ('ii', C.c_int32), ('fi', C.c_int32), ('if_', C.c_int32), ('ff', C.c_int32),
                ('mult', C.c_int32), ('reserved', C.c_int32), ('strength', C.c_double)]


class StbStats(C.Structure):
    _fields_ = [('n_steps', C.c_int64), ('n_rejected', C.c_int64), ('n_newton_fail', C.c_int64),
                ('n_rhs', C.c_int64), ('n_jv', C.c_int64), ('n_lin', C.c_int64),
                ('n_out', C.c_int64), ('status', C.c_int32), ('last_order', C.c_int32),
                ('t_reached', C.c_double), ('last_h', C.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


EXPORTS = ['stb_last_error', 'stb_device_count', 'stb_describe', 'stb_system_create',
           'stb_system_destroy', 'stb_system_info', 'stb_ground_states', 'stb_set_batch',
           'stb_set_blocks', 'stb_set_strengths', 'stb_num_interactions',
           'stb_export_interactions', 'stb_rhs', 'stb_jv', 'stb_rhs_dev', 'stb_jv_dev',
           'stb_to_planes', 'stb_from_planes', 'stb_integrate', 'stb_integrate_dev', 'stb_integrate_dynamics', 'stb_integrate_dynamics_dev', 'stb_last_timing_dynamics', 'stb_set_option', 'stb_get_option', 'stb_set_experiment', 'stb_objective', 'stb_stream', 'stb_launch_count', 'stb_release_cached_memory',
           'stb_set_coop', 'stb_coop_count', 'stb_export_coop',
           'stb_last_timing', 'stb_last_profile']

_lib = None


class StbError(RuntimeError):
    """A C-ABI call failed; `.code` is the STB_E* value."""

    def __init__(self, code, msg):
        super().__init__('%s (code %d)' % (msg, code))
        self.code = code


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('simetuc_b200: %s not built; run `python -c "import __graft_entry__ as g; '
                          'g.build()"` or `make -C simetuc_b200/csrc` (no CPU fallback exists)'
                          % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    dp = C.POINTER(C.c_double)
    lib.stb_last_error.restype = C.c_char_p
    lib.stb_device_count.restype = C.c_int
    lib.stb_describe.argtypes = [C.c_char_p, C.c_int64]
    lib.stb_system_create.restype = P
    lib.stb_system_create.argtypes = [C.c_int32, C.c_int64, P, C.c_int32, C.c_int32, P, P, P,
                                      C.c_double, C.c_int32, P]
    lib.stb_system_destroy.argtypes = [P]
    lib.stb_system_destroy.restype = None
    lib.stb_system_info.argtypes = [P, C.POINTER(C.c_int64)]
    lib.stb_ground_states.argtypes = [P, P, P, P]
    lib.stb_set_batch.argtypes = [P, C.c_int32]
    lib.stb_set_blocks.argtypes = [P, C.c_int32, C.c_int32, P, P]
    lib.stb_set_strengths.argtypes = [P, C.c_int32, P]
    lib.stb_num_interactions.argtypes = [P, C.c_int32]
    lib.stb_num_interactions.restype = C.c_int64
    lib.stb_export_interactions.argtypes = [P, C.c_int32, P, P, P, P]
    lib.stb_rhs.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, P, P]
    lib.stb_jv.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, P, P, P]
    lib.stb_rhs_dev.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, P, P, C.c_int32,
                                C.POINTER(C.c_float)]
    lib.stb_jv_dev.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, P, P, P, C.c_int32,
                               C.POINTER(C.c_float)]
    lib.stb_to_planes.argtypes = [P, C.c_int32, P, P]
    lib.stb_from_planes.argtypes = [P, C.c_int32, P, P]
    lib.stb_integrate.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, P, P,
                                  C.c_double, C.c_double, C.c_int64, P, P, P]
    lib.stb_integrate_dev.argtypes = lib.stb_integrate.argtypes
    lib.stb_set_option.argtypes = [P, C.c_char_p, C.c_double]
    lib.stb_get_option.argtypes = [P, C.c_char_p, C.POINTER(C.c_double)]
    lib.stb_set_experiment.argtypes = [P, C.c_int32, C.c_int64, P, P, C.c_double]
    lib.stb_objective.argtypes = [P, C.c_int32, P]
    lib.stb_last_timing_dynamics.argtypes = [P, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.stb_integrate_dynamics.argtypes = [P, C.c_int32, C.c_int32, C.c_int64, P, C.c_int64, P, P,
                                           C.c_double, C.c_double, C.c_int64, C.c_int64, P, P, P, P, P]
    lib.stb_integrate_dynamics_dev.argtypes = lib.stb_integrate_dynamics.argtypes
    lib.stb_stream.argtypes = [P]
    lib.stb_stream.restype = C.c_void_p
    lib.stb_launch_count.restype = C.c_int64
    lib.stb_release_cached_memory.restype = C.c_int64
    lib.stb_set_coop.argtypes = [P, P, P, P, P, C.c_int32, C.c_double, C.c_double]
    lib.stb_coop_count.argtypes = [P]
    lib.stb_coop_count.restype = C.c_int64
    lib.stb_export_coop.argtypes = [P, P, P, P]
    lib.stb_last_timing.argtypes = [P, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    lib.stb_last_profile.argtypes = [P, C.POINTER(C.c_double)]
    _lib = lib
    return lib


def last_error():
    return load().stb_last_error().decode(errors='replace')


def check(rc):
    if rc != 0:
        raise StbError(rc, last_error())


def require_gpu():
    lib = load()
    if lib.stb_device_count() < 1:
        raise RuntimeError('simetuc_b200: no CUDA device visible; the rate-equation path has no '
                           'CPU fallback')
    return lib


def ptr(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a
