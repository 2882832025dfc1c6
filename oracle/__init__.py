"""TEST INFRASTRUCTURE -- CPU oracle for the fused PDE-residual path.

Nothing under ``oracle/`` is shipped or imported by ``pinnacle_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs use it.
See ``oracle/autograd_ref.py`` (reference algorithm: nested reverse-mode autograd) and
``oracle/jets_spec.py`` (fp64 executable specification of the Taylor-jet formulation).
"""
