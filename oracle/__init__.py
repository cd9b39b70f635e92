"""CPU oracle for the sGDML hot paths -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain NumPy/SciPy restatement of the reference algorithm (stefanch/sGDML v1.0.3,
commit a6ae5e8) for the two paths in SURVEY.md section 8: (a) descriptor ->
Hessian-kernel matrix K -> Cholesky solve -> model, (b) energy/force prediction.
Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker / the
reported CPU baseline.  Nothing under ``sgdml_b200/`` imports it.

Parity pin: the reference ships no tests, fixtures or golden vectors,
so this restatement is pinned against the reference ITSELF, imported in the build
container from a writable copy of /root/reference (``baseline/_ref``):
``tests/golden/make_golden.py`` runs the unmodified reference on seeded synthetic
inputs and freezes its outputs under ``tests/golden/*.npz``;
``tests/test_oracle_vs_golden.py`` checks every oracle function against them.
"""

from . import desc, assemble, solve, predict, train, iterative  # noqa: F401
