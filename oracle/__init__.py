"""CPU oracle for the IRN pseudo-label hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``irn_b200/`` imports this package; the only
callers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker / the CPU baseline, never as the
product path.

What it is: a restatement, in numpy / torch-CPU fp32, of the reference's algorithm for the
path SURVEY.md section 8 names.  Every function cites the reference file:line it follows.

Parity pinning: the reference has no tests, golden vectors or fixtures of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
imported unmodified from ``/root/reference`` in the build container by
``tests/golden/make_golden.py`` (committed, together with the fixtures it wrote under
``tests/golden/``).  ``tests/test_oracle_golden.py`` checks every oracle function against
those fixtures.  Third-party boundaries the reference leans on (torch conv/GN/interpolate,
skimage.measure.label, PIL bicubic) are pinned only through those runs: skimage is absent
here, so connected-component labelling is "parity unpinned" beyond scipy.ndimage.label.
"""
