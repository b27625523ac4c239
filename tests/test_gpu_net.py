"""GPU: BASELINE.json configs[4] -- a random-init ZebraPose network's bf16 forward on 256 x 256 crops feeding the
device-side decode + RANSAC without a host copy (logits of a random-init net are noise, so this checks
decode parity and the plumbing, not pose accuracy).

The network body is torch / cuDNN (workloads/net.py, pinned against the reference model on the CPU by
tests/test_net_feeder.py); the path under test starts at its last activations: `Engine.head_decode` (fused conv_1x1_4 +
threshold + pack + emit) and `head_pose_batch`.

Floating-point tolerance: a bit is (sum_c bf16(w) bf16(x) accumulated in fp32 + bias) > 0.  The rigorous bound of an
fp32 accumulation of n = 320 terms is n * 2^-23 * sum|w x| (3.8e-5 sum|w x|); bits may differ from the float64 sum of the same
bf16 values only where |logit| <= 1e-4 * (sum|w x| + |bias|), everywhere else they must be equal."""
import os

import numpy as np
import pytest
import torch

from workloads import net as znet
from workloads import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_net_v1.npz"))


@pytest.fixture(scope="module")
def eng():
    import zebrapose_b200 as zp
    e = zp.Engine(0)
    tab, _, _ = synth.make_dict(16, seed=5, radius=60.0, missing_frac=0.1)
    e.upload_dict(1, tab, n_bits=16, ignore_bit=0, nonexist="zero")
    return e


def _cl(t):
    return t.permute(0, 2, 3, 1).is_contiguous()


def test_feeder_on_the_device_reproduces_reference_logits():
    """float32 body on the GPU (TF32 off) against the reference model's logits: the device feeder is the reference graph.
    Tolerance 2e-4 absolute on logits of std 0.06 (cuDNN may pick Winograd / FFT algorithms for the float32 3x3 layers)"""
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = znet.build(seed=0, device="cuda")
        img = znet.images(2, seed=0, device="cuda")
        with torch.no_grad():
            lg = net.logits(img).float().cpu().numpy()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    np.testing.assert_allclose(lg[:, :, ::8, ::8], G["logits_sub"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(lg[0, :, 77, :], G["logits_row"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("fold", [True, False])
def test_config5_bf16_activations_through_fused_head(eng, fold):
    B, S = 6, 128
    net = znet.build(seed=0, device="cuda", dtype=torch.bfloat16, fold=fold)
    img = znet.images(B, seed=0, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        x, xs = net(img)
    assert x.shape == (B, 256, S, S) and xs.shape == (B, 64, S, S) and x.dtype == torch.bfloat16
    assert _cl(x) and _cl(xs), "cuDNN hands the activations over channels_last: consumed in place, no copy"
    W = net.tail.weight.detach().float().reshape(17, 320)           # bf16 values
    bias = net.tail.bias.detach().float()
    eng.upload_head(W, bias)
    bb = np.tile(np.array([[100.0, 60.0, 180.0, 180.0]]), (B, 1))
    corr, counts, codes = eng.head_decode(x, xs, bb, obj_default=1, return_codes=True)

    feats = torch.cat([x, xs], 1).double()
    ref = torch.einsum("oc,bchw->bohw", W.double(), feats) + bias.double().view(1, -1, 1, 1)
    bound = torch.einsum("oc,bchw->bohw", W.double().abs(), feats.abs()) + bias.double().abs().view(1, -1, 1, 1)
    sure = ref.abs() > REL_TOL * bound
    assert sure.float().mean().item() > 0.98
    w = (2 ** torch.arange(15, -1, -1, device="cuda")).view(1, 16, 1, 1)
    want_codes = ((ref[:, 1:] > 0).long() * w).sum(1)
    bit_diff = (((codes.long() ^ want_codes).unsqueeze(1) // w) % 2).bool()
    assert not (bit_diff & sure[:, 1:]).any(), "a code bit differs where the float64 logit is not within the fp32 bound of zero"
    want_mask = ref[:, 0] > 0
    lo = (want_mask & sure[:, 0]).flatten(1).sum(1)
    hi = (want_mask | ~sure[:, 0]).flatten(1).sum(1)
    c = counts.long()
    assert bool(((c >= lo) & (c <= hi)).all()) and int(c.min()) > 1000       # noise logits: most pixels are "masked"

    # crops whose every mask pixel is certain: the correspondence list is the unfused path's on the materialised logits
    lg = ref.float().contiguous()
    corr2, counts2, codes2 = eng.decode(lg, bb, None, obj_default=1, return_codes=True)
    certain = sure.flatten(1).all(1).cpu().numpy()
    for i in np.nonzero(certain)[0]:
        n = int(counts[i])
        assert n == int(counts2[i]) and torch.equal(corr[i, :, :n], corr2[i, :, :n])
    d2 = (((codes.long() ^ codes2.long()).unsqueeze(1) // w) % 2).bool()
    assert not (d2 & sure[:, 1:]).any()

    # whole chain: activations -> poses on the device; noise correspondences give no meaningful pose, only a valid status,
    # and the chain is deterministic (same bits, same seeded samples -> same answer)
    K = np.tile(synth.YCBV_K.reshape(1, 9), (B, 1))
    p1, n1, s1 = eng.head_pose_batch(x, xs, bb, K, obj_default=1)
    p2, n2, s2 = eng.head_pose_batch(x, xs, bb, K, obj_default=1)
    torch.cuda.synchronize()
    assert bool(torch.isin(s1, torch.tensor([0, 3], device="cuda", dtype=s1.dtype)).all())
    ok = s1 == 0
    assert bool(torch.isfinite(p1[ok]).all())
    assert torch.equal(s1, s2) and torch.equal(n1, n2) and torch.equal(p1[ok], p2[ok])


def test_config5_shards_are_independent(eng):
    """1024 crops over 8 ranks = 128 per rank; here: a shard's result does not depend on what else is in the batch"""
    import zebrapose_b200 as zp
    B = 8
    net = znet.build(seed=0, device="cuda", dtype=torch.bfloat16, fold=True)
    img = znet.images(B, seed=3, device="cuda", dtype=torch.bfloat16)
    eng.upload_head(net.tail.weight.detach().float().reshape(17, 320), net.tail.bias.detach().float())
    bb = torch.tensor([[100.0, 60.0, 180.0, 180.0]], device="cuda", dtype=torch.float64).repeat(B, 1)
    K = torch.from_numpy(np.tile(synth.YCBV_K.reshape(1, 9), (B, 1))).cuda()
    with torch.no_grad():
        x, xs = net(img)
    whole = eng.head_pose_batch(x, xs, bb, K, obj_default=1)
    for r in range(2):
        lo, hi = zp.shard_range(B, r, 2)
        xr = x[lo:hi].contiguous(memory_format=torch.channels_last)
        xsr = xs[lo:hi].contiguous(memory_format=torch.channels_last)
        part = eng.head_pose_batch(xr, xsr, bb[lo:hi], K[lo:hi], obj_default=1)
        assert torch.equal(part[2], whole[2][lo:hi]) and torch.equal(part[1], whole[1][lo:hi])
        ok = part[2] == 0
        assert torch.equal(part[0][ok], whole[0][lo:hi][ok])
