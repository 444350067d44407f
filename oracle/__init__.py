"""CPU oracle for the flow hot path (TEST INFRASTRUCTURE -- not a product path).

A numpy/scipy restatement of what nschloe/flow asks DOLFIN/FFC/PETSc to compute
for ``flow.navier_stokes.{Chorin,IPCS,Rotational}.step``, ``flow.heat.Heat`` and
``flow.stokes.solve``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; ``flow_b200`` never does.

PARITY UNPINNED at the bit level: the reference's arithmetic lives in DOLFIN /
FFC / PETSc / hypre, none of which is vendored, pinned or installable here
(SURVEY.md section 8c), and the reference tree holds no golden vectors that can be
reproduced without gmsh + ``materials``.  What *is* pinned, and checked in
``tests/test_oracle_*.py``, are the reference's own acceptance thresholds:
manufactured-solution temporal orders (tests/test_navier_stokes.py:386-445),
Stokes spatial order (tests/test_stokes.py:105-117) and the hydrostatic
invariant (tests/test_sealed_box.py:134-141).
"""
