"""CPU oracle for the mc3d hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference's algorithm for the path
BASELINE.json:north_star names (DLT triangulation, heatmap decode,
refinement loss/gradient/Adam).  Every function cites the reference
file:line it follows (paths relative to the upstream repo).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and
only as the checker / the CPU baseline.  The product package
(``multi-camera_3d_pose_estimation_b200``) never imports it and has no CPU
fallback: it raises when the CUDA library is missing.

Parity pinning: the oracle is pinned by ``tests/golden/*.npz`` -- outputs of
the UNMODIFIED reference functions, produced in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` with
stubs for the unused matplotlib / mmpose imports) -- see
``tests/test_oracle_golden.py``.  The argmax decode (third-party mmpose,
unvendored and unpinned upstream) is the one piece that is "parity unpinned";
``oracle/decode.py`` says so where it is defined.
"""
