"""ORACLE -- CPU (numpy) restatement of the reference's CPU tensor path.

Test infrastructure only: imported by ``tests/``, ``__graft_entry__.smoke()``
and the cpu_baseline / ``--impl reference`` legs of ``bench.py``; the product
package ``lightgrad_b200`` never imports it.  See ``np_ops.py`` for the
per-rule citations and ``tests/golden/make_golden.py`` for how it is pinned to
outputs of the real reference.
"""
from .cpu import CpuTensor  # noqa: F401
