"""pytest plugin (build container only): run the REFERENCE's own known-answer tests with the
oracle restatement substituted for `precalculate.setup_microscopic_eqs`.

If `simetuc/test/test_setup/test_setup.py::test_lattice_*` ("THESE RESULTS HAVE BEEN CHECKED BY
HAND", `test_setup.py:22-23`) pass with `oracle.setup_eqs.setup_microscopic_eqs` in place of the
reference builder, the oracle reproduces every hand-checked matrix, pair list and Jacobian
index list the reference pins.  TEST INFRASTRUCTURE.
"""
import numpy as np

from oracle import ref_import

ref_import.install_stubs()
if not hasattr(np, 'Inf'):          # numpy>=2 dropped the alias used at precalculate.py:590
    np.Inf = np.inf
_ns = ref_import.load()

import oracle.setup_eqs as _setup  # noqa: E402

_ns.precalculate.setup_microscopic_eqs = _setup.setup_microscopic_eqs
