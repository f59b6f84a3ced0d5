"""CPU oracle for the GO-MELT multilevel explicit thermal time-step.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``gomelt_b200/``) never does, and fails
loudly when its CUDA library is missing.

What it is: a float32 NumPy restatement of the reference's algorithm *as written*
(``/root/reference/go_melt/computeFunctions.py``, ``go_melt.py``, ``createPath.py``):
per-element gather -> 8x8 ``diag(Me) - Ke`` apply -> ordered f32 scatter-add, materialised
``(n_fine_elem, 8, 8)`` transfer operators, full-field interpolation for face BCs.  Every
function cites the reference ``file:line`` it follows (``cF`` = computeFunctions.py,
``gm`` = go_melt.py, ``cP`` = createPath.py).

Parity pin status
-----------------
The reference has no tests, golden vectors or published numbers, and JAX (its array
engine, pinned ``jax[cuda12_pip]==0.4.16``) is not installable here, so XLA's own float
ordering is **unpinned**.  What *is* pinned (``tests/golden/``):

* the reference's own source, executed unmodified on small meshes through a NumPy-backed
  stand-in for the ``jax`` API (``tests/golden/jax_numpy_shim.py`` +
  ``tests/golden/make_golden.py``), produced the committed ``*.npz`` fixtures the oracle is
  checked against (``tests/test_oracle_golden.py``);
* ``createPath.parsingGcode`` runs as-is in this container; its toolpath for
  ``examples/example.gcode`` is a committed fixture;
* analytic known-answer tests that follow from the reference's formulas (SURVEY.md section 4,
  K1-K7).

dtype: ``oracle.config.FDT`` (np.float32 = the reference dtype; tests may flip it to
float64 to estimate round-off).
"""
from . import config  # noqa: F401
