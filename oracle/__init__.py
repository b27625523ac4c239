"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy / cv2) of the reference's post-network pose path
(lyltc1/ZebraPose, zebrapose/binary_code_helper/* + zebrapose/common_ops.py +
the cv2.solvePnPRansac call in CNN_output_to_pose.py:155-158).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and only as the checker or
as the timed CPU baseline.  The product package `zebrapose_b200` never imports
it and fails loudly when its CUDA library is missing.

Parity pin: the reference has no tests / golden vectors for this path
(SURVEY.md section 4).  The restatement is pinned instead against outputs of the
*reference functions themselves*, imported from /root/reference in the build
container by `tests/golden/make_golden.py`; the resulting fixtures are
committed under `tests/golden/` and checked by `tests/test_oracle_golden.py`.
The RANSAC/EPnP arithmetic lives in OpenCV (un-vendored; opencv-python-headless
4.13.0.92 in this image): `oracle.cvransac` is a control-flow emulation that was
verified to return identical inlier sets, and `oracle.epnp` a float64
restatement of OpenCV's EPnP that agrees with `cv2.solvePnP(SOLVEPNP_EPNP)` to
~1e-5 deg for n >= 6 (for 4/5-point minimal sets the null space of M^T M is
degenerate and cv2's own answer depends on rounding noise -- "parity unpinned"
for those hypotheses).
"""
