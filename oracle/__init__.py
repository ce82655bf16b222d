"""CPU oracle for the Kriging-over-Saltelli -> Sobol' hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``batman_b200/`` may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` use it, and there only as the
checker (or as the timed CPU arm), never as the product.

Layout
------
``gp.py``        numpy restatement of the scikit-learn GP arithmetic that
                 ``batman/surrogate/kriging.py`` drives (kernels, predict, LML).
``gp_hp.py``     the same formulas in ``np.longdouble``: arbiter at 1e-10.
``sk_reference`` the reference call sequence of ``kriging.py`` on the installed
                 scikit-learn (per-point ``multi_eval`` loop and batched).
``saltelli.py``  numpy restatement of the OpenTURNS 1.15 Saltelli design layout
                 and estimator, plus batman's own aggregation glue (uq.py).
``philox.py``    counter-based design generator shared (bit-exactly) with the
                 CUDA path, for designs too large to store.
``pod.py``       ``Pod.inverse_transform`` (pod/pod.py:219-228).
``functions.py`` the analytical test functions used as training-data makers.
``ref_shim.py``  imports the real reference from /root/reference (this
                 container only) to pin the restatements; never used at GPU
                 test time.

Parity status: the GP half is pinned against scikit-learn (installed, the very
library the reference calls) and against the reference's own ``Kriging`` via
``tests/golden``.  The Saltelli half is **parity unpinned**: OpenTURNS is not
installable here, so the estimator conventions are a documented restatement
anchored on analytic Sobol' indices only (see saltelli.py header).
"""
