"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference algorithm for MaskPlanner's point-cloud hot path (PointNet++
set abstraction + chamfer set loss).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import anything from here, and only
as the checker or the timed CPU baseline.  ``maskplanner_b200/`` never imports this package.

Parity pin
----------
* Encoder side (FPS, ball query, gather/grouping, set-abstraction module): pinned.  The restatement
  is checked bit-for-bit against the reference's own code imported from ``/root/reference`` by
  ``oracle/make_golden.py`` and against the fixtures that script froze under ``tests/golden/``.
* Chamfer wrapper logic (padding scan, masking, reductions, direction selection): pinned the same
  way -- the reference's ``pytorch3d_chamfer.py`` is imported unmodified over a shim.
* The nearest-neighbour arithmetic underneath the wrapper lives in ``pytorch3d`` (v0.7.0 / v0.7.2 /
  commit c292c71c, pinned only in the reference README, not vendored, not installable offline):
  **parity unpinned** at that boundary.  ``knn_points`` below restates pytorch3d's published
  semantics (squared L2, direct form, lengths honoured, lowest-index tie-break).

Modules
-------
``c_oracle``      ctypes view of ``oracle/c/mp_oracle.c`` (bit-exact index arithmetic, OpenMP).
``torch_oracle``  the same algorithm in plain CPU torch ops, following the reference's op order.
``ref_loader``    imports the real reference from ``/root/reference`` (this container only).
``make_golden``   writes ``tests/golden/*.npz`` from the real reference.
"""
