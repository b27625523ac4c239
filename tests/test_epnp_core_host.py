"""CPU: the *device* EPnP core (zebrapose_b200/csrc/zp_epnp.cuh) compiled for the host by nvcc and checked against
cv2.solvePnP(SOLVEPNP_EPNP) -- the same source the CUDA kernels run, so solver parity is testable without a GPU."""
import os
import subprocess

import numpy as np
import pytest

from oracle import cvransac, metrics
from workloads import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("native") / "epnp_host")
    src = os.path.join(ROOT, "tests", "native", "epnp_host.cu")
    r = subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _solve(exe, problems, K, mode=1):
    txt = []
    for pw, uv in problems:
        txt.append("%d %d" % (len(pw), mode))
        txt.append(" ".join("%.17g" % v for v in K.ravel()))
        txt += [" ".join("%.17g" % v for v in list(p) + list(q)) for p, q in zip(pw, uv)]
    out = subprocess.run([exe], input="\n".join(txt) + "\n", capture_output=True, text=True, check=True).stdout
    a = np.array([[float(v) for v in line.split()] for line in out.strip().split("\n")])
    return a[:, :9].reshape(-1, 3, 3), a[:, 9:]


@pytest.mark.parametrize("n,noise", [(6, 0.0), (6, 1.0), (6, 30.0), (8, 30.0), (50, 30.0), (2000, 2.0)])
def test_device_core_matches_cv2(harness, n, noise):
    rng = np.random.default_rng(int(n * 100 + noise))
    K = synth.YCBV_K if n == 8 else synth.LM_K
    problems, ref = [], []
    for _ in range(12):
        pw = (rng.normal(size=(n, 3)) * 40).astype(np.float32)
        R, t = synth.random_pose(rng)
        P = (R @ pw.T).T + t
        uv = (K @ P.T).T
        uv = np.trunc(uv[:, :2] / uv[:, 2:] + rng.normal(size=(n, 2)) * noise).astype(np.float32)
        problems.append((pw, uv))
        ref.append(cvransac.cv2_solver(pw, uv, K))
    Rs, ts = _solve(harness, problems, K)
    for (Rc, tc), R, t in zip(ref, Rs, ts):
        assert metrics.rot_err_deg(Rc, R) < 5e-4
        assert metrics.trans_err(tc, t) < 5e-3
    # the split final solve's arithmetic (raw moments relative to the first point + contractions): the same pose to rounding
    Rm, tm = _solve(harness, problems, K, mode=2)
    scale = 1e-7 if noise >= 30.0 and n <= 8 else 1e-9       # tiny noisy samples are ill-conditioned: rounding is amplified
    assert np.abs(Rm - Rs).max() < scale and np.abs(tm - ts).max() < scale * 1e3, (np.abs(Rm - Rs).max(), np.abs(tm - ts).max())
