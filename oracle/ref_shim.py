"""Import the real reference (tupui/batman 1.9) from /root/reference.

TEST INFRASTRUCTURE ONLY, and only in the build container: /root/reference does
not exist on the GPU box, so nothing at GPU test time, in ``smoke()`` or in
``bench.py`` may call this.  It is used by tests/golden/gen_golden.py to pin
the restatements under ``oracle/`` against the reference's own classes.

``import batman`` fails as shipped in this image: openturns,
pathos and matplotlib are absent, and ``scipy.integrate.simps`` was removed.
The shim pre-seeds ``sys.modules`` with stand-ins for exactly those, which the
Kriging / SurrogateModel / Pod / analytical-function code never executes.
"""
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'batman'))


def import_batman():
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    if 'batman' in sys.modules and getattr(sys.modules['batman'], '__shimmed__', False):
        return sys.modules['batman']
    for name in ['openturns', 'openturns.viewer', 'matplotlib', 'matplotlib.pyplot',
                 'matplotlib.backends', 'matplotlib.backends.backend_pdf', 'matplotlib.colors',
                 'matplotlib.cm', 'matplotlib.animation', 'matplotlib.ticker', 'matplotlib.patches',
                 'matplotlib.path', 'matplotlib.tri', 'matplotlib.lines', 'matplotlib.collections',
                 'mpl_toolkits', 'mpl_toolkits.mplot3d', 'batman.visualization']:
        sys.modules.setdefault(name, mock.MagicMock())
    import multiprocess
    import multiprocess.pool
    pathos = types.ModuleType('pathos')
    pmp = types.ModuleType('pathos.multiprocessing')
    pmp.Pool = multiprocess.pool.Pool
    pmp.cpu_count = multiprocess.cpu_count
    pathos.multiprocessing = pmp
    sys.modules.setdefault('pathos', pathos)
    sys.modules.setdefault('pathos.multiprocessing', pmp)
    import scipy.integrate
    if not hasattr(scipy.integrate, 'simps'):
        scipy.integrate.simps = scipy.integrate.simpson
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import batman
    batman.__shimmed__ = True
    return batman
