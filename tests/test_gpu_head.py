"""GPU: the fused network tail (zp_upload_head + zp_head_decode: 1x1 convolution on tcgen05 tensor cores -> threshold ->
bit pack -> correspondence lists) against a plain PyTorch reference of the same convolution followed by the CPU oracle.

Floating-point tolerance: a bit is (sum_c bf16(w) * bf16(x) accumulated in fp32 + bias) > 0.  Tensor-core and reference
summation orders differ, so a bit may legitimately differ only where the reference logit (float64 sum of the same bf16
products) is within LOGIT_TOL = 2e-3 of zero (products are O(1), 320 terms); everywhere else codes and mask must be
bit-exact, and on the structured inputs (|logit| >= 0.9) the correspondence lists must equal the oracle's exactly."""
import numpy as np
import pytest
import torch

from oracle import decode as odec
from workloads import synth

pytestmark = pytest.mark.gpu
LOGIT_TOL = 2e-3


@pytest.fixture(scope="module")
def eng():
    import zebrapose_b200 as zp
    return zp.Engine(0)


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _features_for(logits, W, bias):
    """activations x (bf16, channels_last) whose 1x1 convolution with W (+ bias) reproduces `logits` up to bf16 rounding:
    the minimum-norm solution x = pinv(W) (logits - bias) per pixel"""
    B, Co, S, _ = logits.shape
    L = torch.from_numpy(logits).cuda().double().permute(0, 2, 3, 1).reshape(-1, Co) - bias.double()
    x = (L @ torch.linalg.pinv(W.double()).T).float()                   # [B*S*S, c_in]
    x = x.reshape(B, S, S, -1).permute(0, 3, 1, 2)                      # logical NCHW over channels-last memory
    return x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def _ref_logits(x, x_skip, W, bias):
    """float64 sum of the SAME bf16 values the kernel multiplies"""
    xs = x if x_skip is None else torch.cat([x, x_skip], 1)
    return torch.einsum("oc,bchw->bohw", W.double(), xs.double()) + bias.double().view(1, -1, 1, 1)


def _expected_codes(ref, mask_ch, bit0_ch, nb):
    bits = (ref[:, bit0_ch:bit0_ch + nb] > 0).long()
    w = (2 ** torch.arange(nb - 1, -1, -1, device=ref.device)).view(1, nb, 1, 1)
    return (bits * w).sum(1), ref[:, mask_ch] > 0


@pytest.mark.parametrize("c1,c2", [(256, 64), (320, 0), (64, 64)])
def test_head_structured_logits_match_oracle_exactly(eng, c1, c2):
    S, B = 128, 5
    tab, nrm, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.1)
    eng.upload_dict(0, tab, n_bits=16, ignore_bit=0, nonexist="zero")
    crops = [synth.make_crop(tab, nrm, 9000 + i, S=S) for i in range(B)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    bboxes = np.stack([c["bbox"] for c in crops])
    g = torch.Generator(device="cpu").manual_seed(1)
    W = _bf16_round(torch.randn(17, c1 + c2, generator=g) * 0.3).cuda()
    bias = (torch.randn(17, generator=g) * 0.2).cuda()
    xall = _features_for(logits, W, bias)
    x = xall[:, :c1].contiguous(memory_format=torch.channels_last)
    xs = xall[:, c1:].contiguous(memory_format=torch.channels_last) if c2 else None
    ref = _ref_logits(x, xs, W, bias)
    assert ref.abs().min().item() > 0.5            # precondition: no logit near zero, so the result is unambiguous
    assert torch.equal(ref > 0, torch.from_numpy(logits).cuda() > 0)
    eng.upload_head(W, bias)
    corr, counts, codes = eng.head_decode(x, xs, bboxes, return_codes=True)
    torch.cuda.synchronize()
    want_codes, want_mask = _expected_codes(ref, 0, 1, 16)
    assert torch.equal(codes.long(), want_codes)
    corr, counts = corr.cpu().numpy(), counts.cpu().numpy()
    for i, c in enumerate(crops):
        mask = odec.threshold_logits(logits[i, 0]).astype(np.uint8)
        code = odec.threshold_logits(logits[i, 1:]).transpose(1, 2, 0)
        uv, xyz, _ = odec.decode_crop(mask, code, c["bbox"], S, tab)
        n = counts[i]
        assert n == len(uv)
        assert np.array_equal(corr[i, 0:2, :n].T, uv)
        assert np.array_equal(corr[i, 2:5, :n].T.view(np.uint32), xyz.view(np.uint32))
    # the unfused path on the materialised logits gives the identical lists
    lg = ref.float().contiguous()
    corr2, counts2 = eng.decode(lg, bboxes)
    assert torch.equal(counts2.cpu(), torch.from_numpy(counts))
    for i in range(B):
        assert torch.equal(corr2[i, :, :counts[i]].cpu(), torch.from_numpy(corr[i, :, :counts[i]]))


@pytest.mark.parametrize("S,B,mask_ch,bit0_ch,n_out,k", [(128, 3, 0, 1, 17, 0), (64, 7, 0, 2, 18, 0), (128, 2, 0, 1, 17, 4),
                                                         (32, 9, 1, 2, 32, 3)])
def test_head_random_features_bits_within_tolerance(eng, S, B, mask_ch, bit0_ch, n_out, k):
    """noise activations (random-init network, BASELINE config 5): every bit whose reference logit is not within
    LOGIT_TOL of zero must match; layouts: 17-channel v1, 18-channel v2 (mask, entire mask, bits), ignore_bit, 32 outputs"""
    g = torch.Generator(device="cpu").manual_seed(S + B)
    c1, c2 = 256, 64
    W = _bf16_round(torch.randn(n_out, c1 + c2, generator=g) * 0.1).cuda()
    bias = (torch.randn(n_out, generator=g) * 0.1).cuda()
    x = torch.randn(B, c1, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xs = torch.randn(B, c2, S, S, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    tab, _, _ = synth.make_dict(16, seed=5, radius=60.0, missing_frac=0.0)
    eng.upload_dict(1, tab, n_bits=16, ignore_bit=k, nonexist="zero")
    eng.upload_head(W, bias)
    nb = 16 - k
    bb = np.tile(np.array([[10.0, 20.0, 200.0, 200.0]]), (B, 1))
    corr, counts, codes = eng.head_decode(x, xs, bb, obj_default=1, mask_ch=mask_ch, bit0_ch=bit0_ch, n_bits=16, ignore_bit=k,
                                          return_codes=True)
    ref = _ref_logits(x, xs, W, bias)
    want_codes, want_mask = _expected_codes(ref, mask_ch, bit0_ch, nb)
    sure = (ref[:, bit0_ch:bit0_ch + nb].abs() > LOGIT_TOL)
    diff = codes.long() ^ want_codes
    w = (2 ** torch.arange(nb - 1, -1, -1, device=ref.device)).view(1, nb, 1, 1)
    bit_diff = ((diff.unsqueeze(1) // w) % 2).bool()
    assert not (bit_diff & sure).any(), "a bit differs where the reference logit is not near zero"
    assert bit_diff.sum().item() <= (~sure).sum().item()
    assert sure.float().mean().item() > 0.99
    # mask: counts must equal the number of positive mask logits up to the near-zero ones
    sure_m = ref[:, mask_ch].abs() > LOGIT_TOL
    lo = (want_mask & sure_m).flatten(1).sum(1).cpu()
    hi = (want_mask | ~sure_m).flatten(1).sum(1).cpu()
    c = counts.cpu().long()
    assert bool(((c >= lo) & (c <= hi)).all())
    # the fused result equals the unfused path run on fp32 logits computed by torch from the same bf16 values wherever sure
    lg = ref.float().contiguous()
    corr2, counts2, codes2 = eng.decode(lg, bb, None, obj_default=1, mask_ch=mask_ch, bit0_ch=bit0_ch, n_bits=16,
                                        ignore_bit=k, return_codes=True)
    d2 = ((((codes.long() ^ codes2.long()).unsqueeze(1)) // w) % 2).bool()
    assert not (d2 & sure).any()


def test_head_argument_checks(eng):
    import zebrapose_b200 as zp
    W = torch.randn(17, 320)
    eng.upload_head(W, None)
    x = torch.zeros(1, 256, 128, 128, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with pytest.raises(zp.ZpError):
        eng.head_decode(x, None, np.zeros((1, 4)))                      # 256 != 320 channels
    with pytest.raises(ValueError):
        eng.head_decode(x.contiguous(), None, np.zeros((1, 4)))         # NCHW-contiguous, not channels_last
    with pytest.raises(zp.ZpError):
        eng.upload_head(torch.randn(40, 320))                            # more than 32 outputs
    with pytest.raises(zp.ZpError):
        eng.upload_head(torch.randn(17, 100))                            # c_in not a multiple of 64


def test_head_pose_batch_equals_logit_path(eng):
    """network activations -> poses in one call == the same crops through decode_and_pose_batch on materialised logits"""
    S, B = 128, 4
    tab, nrm, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.0)
    eng.upload_dict(0, tab, n_bits=16, ignore_bit=0, nonexist="zero")
    crops = [synth.make_crop(tab, nrm, 9100 + i, S=S) for i in range(B)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    bboxes = np.stack([c["bbox"] for c in crops])
    Ks = np.stack([c["K"] for c in crops]).reshape(B, 9)
    g = torch.Generator(device="cpu").manual_seed(2)
    W = _bf16_round(torch.randn(17, 320, generator=g) * 0.3).cuda()
    bias = (torch.randn(17, generator=g) * 0.2).cuda()
    xall = _features_for(logits, W, bias)
    x = xall[:, :256].contiguous(memory_format=torch.channels_last)
    xs = xall[:, 256:].contiguous(memory_format=torch.channels_last)
    eng.upload_head(W, bias)
    p1, n1, s1 = eng.head_pose_batch(x, xs, bboxes, Ks)
    p2, n2, s2 = eng.decode_and_pose_batch(torch.from_numpy(logits).cuda(), bboxes, Ks)
    assert torch.equal(p1, p2) and torch.equal(n1, n2) and torch.equal(s1, s2)
    assert int(s1.sum()) == 0


def test_head_fp32_activations_tf32(eng):
    """float32 channels_last activations take the kind::tf32 tensor-core path: products of the operands' top 10 mantissa
    bits, fp32 accumulation.  Against the float64 convolution of the full-precision values a bit may differ only where
    |logit| <= TF32_TOL = 3e-2 (320 terms of relative error <= 2^-10 on products of magnitude <= ~1); on the structured
    activations (|logit| > 0.5) the correspondence lists equal the oracle's exactly."""
    TF32_TOL = 3e-2
    S, B = 128, 3
    tab, nrm, _ = synth.make_dict(16, seed=3, radius=51.0, missing_frac=0.1)
    eng.upload_dict(0, tab, n_bits=16, ignore_bit=0, nonexist="zero")
    crops = [synth.make_crop(tab, nrm, 9200 + i, S=S) for i in range(B)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    bboxes = np.stack([c["bbox"] for c in crops])
    g = torch.Generator(device="cpu").manual_seed(4)
    W = (torch.randn(17, 320, generator=g) * 0.3).cuda()
    bias = (torch.randn(17, generator=g) * 0.2).cuda()
    B_, Co = logits.shape[0], logits.shape[1]
    L = torch.from_numpy(logits).cuda().double().permute(0, 2, 3, 1).reshape(-1, Co) - bias.double()
    xall = (L @ torch.linalg.pinv(W.double()).T).float().reshape(B_, S, S, -1).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    x = xall[:, :256].contiguous(memory_format=torch.channels_last)
    xs = xall[:, 256:].contiguous(memory_format=torch.channels_last)
    ref = _ref_logits(x, xs, W, bias)
    assert ref.abs().min().item() > 0.5
    eng.upload_head(W, bias)
    corr, counts, codes = eng.head_decode(x, xs, bboxes, return_codes=True)
    want_codes, _ = _expected_codes(ref, 0, 1, 16)
    assert torch.equal(codes.long(), want_codes)
    corr, counts = corr.cpu().numpy(), counts.cpu().numpy()
    for i, c in enumerate(crops):
        mask = odec.threshold_logits(logits[i, 0]).astype(np.uint8)
        code = odec.threshold_logits(logits[i, 1:]).transpose(1, 2, 0)
        uv, xyz, _ = odec.decode_crop(mask, code, c["bbox"], S, tab)
        n = counts[i]
        assert n == len(uv) and np.array_equal(corr[i, 0:2, :n].T, uv)
        assert np.array_equal(corr[i, 2:5, :n].T.view(np.uint32), xyz.view(np.uint32))
    # noise activations: bits away from zero must match
    bb = np.tile(np.array([[0.0, 0.0, 64.0, 64.0]]), (2, 1))
    Wr = (torch.randn(17, 128, generator=g) * 0.1).cuda()
    xr = torch.randn(2, 128, 64, 64, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    eng.upload_head(Wr, None)
    codes = eng.head_decode(xr[:, :96].contiguous(memory_format=torch.channels_last), xr[:, 96:].contiguous(memory_format=torch.channels_last),
                            bb, return_codes=True)[2]              # 96 + 32: multiples of 32 are enough for fp32
    ref = _ref_logits(xr, None, Wr, torch.zeros(17, device="cuda"))
    want, _ = _expected_codes(ref, 0, 1, 16)
    sure = ref[:, 1:17].abs() > TF32_TOL
    w = (2 ** torch.arange(15, -1, -1, device="cuda")).view(1, 16, 1, 1)
    bit_diff = ((((codes.long() ^ want).unsqueeze(1)) // w) % 2).bool()
    assert not (bit_diff & sure).any()
    assert sure.float().mean().item() > 0.9
