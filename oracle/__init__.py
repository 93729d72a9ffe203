"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the stencil hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and the CPU-baseline / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``stencil_code_b200`` never does and fails loudly without its CUDA shim.

What it is: a restatement of the C and OpenMP-C text that the reference's code generators
(``stencil_code/backend/c.py``, ``backend/omp.py``, ``backend/stencil_backend.py``) build for a
given ``Stencil`` and argument shapes, compiled with ``gcc -O2 -std=c99 [-fopenmp]`` and called
through ctypes with the reference's own entry-point signature
``stencil_kernel(in..., out, float* duration)`` (reference ``stencil_kernel.py:351-367``).

Why a restatement: the reference cannot run here.  Its generators print through the ``ctree``
package (``ctree>=0.95a`` in ``requirements.txt:3``, unpinned, not vendored) and the driver also
imports ``pycl`` and ``hindemith`` (``stencil_kernel.py:24,40,50``); none is installed and there
is no network.  No arithmetic lives in those packages -- ctree only prints the C that
``backend/c.py`` assembles and runs gcc -- so the algorithm is fully determined by the files in
``/root/reference``.  Assumptions about ctree's printing that cannot be checked here: a Python
``int`` prints as a C int literal and a Python ``float`` as a C double literal (``str(value)``),
and the default flags are ``-O2 -std=c99`` (so no FMA contraction on x86-64).

Pinning (parity is PINNED, see ``tests/test_oracle_golden.py``): the oracle is checked against
  * every known-answer value of the reference's own tests for this path
    (``test/test_boundary_handling.py``), and
  * outputs of the UNMODIFIED reference ``Stencil`` class running its ``backend='python'`` path in
    this container (``tests/golden/make_golden.py`` imports it from ``/root/reference`` with inert
    stand-ins for the three missing packages, which that path never calls) -- committed as
    ``tests/golden/*.npz`` for all library kernels x boundary modes.
"""
