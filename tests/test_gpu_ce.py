"""GPU: zp_decode_ce (CE heads of the ablation configs) against the golden digits produced by the reference's own
from_output_to_class_binary_code(..., "CE") (tests/golden/make_golden_ce.py: ties, 1-ulp-apart logits, signed zeros,
saturating softmax; base 2 x 16 digits and base 3 x 8 digits) and the CPU oracle for the correspondence lists."""
import os

import numpy as np
import pytest
import torch

from oracle import decode as odec
from workloads import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    import zebrapose_b200 as zp
    return zp.Engine(0)


@pytest.mark.parametrize("key_in,key_code,base,L", [("ce_in", "ce_code_16", 2, 16), ("ce_in_b3", "ce_code_b3", 3, 8)])
def test_ce_decode_matches_reference(eng, key_in, key_code, base, L):
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_ce_v1.npz"), allow_pickle=False)
    x, digits = g[key_in], g[key_code].astype(np.int64)            # [B, base*L, S, S], [B, L, S, S]
    B, _, S, _ = x.shape
    tab, _, _ = synth.make_dict(16, seed=9, radius=40.0, missing_frac=0.15)
    eng.upload_dict(2, tab, n_bits=16, ignore_bit=0, nonexist="zero")
    rng = np.random.default_rng(0)
    mask = rng.random((B, S, S)) < 0.7
    mask[0, 0, :] = True
    mask[-1] = False                                                # an empty crop
    bb = np.array([[37.0, -12.0, 3 * S + 1, 2 * S + 5]] * B)
    lg = torch.from_numpy(x).cuda()
    corr, counts, codes = eng.decode_ce(lg, bb, base=base, n_digits=L, obj_default=2, digit0_ch=0, ext_mask=mask, return_codes=True)
    w = base ** np.arange(L - 1, -1, -1)
    want_ids = (digits * w[None, :, None, None]).sum(1)
    assert np.array_equal(codes.cpu().numpy().astype(np.int64), want_ids)
    corr, counts = corr.cpu().numpy(), counts.cpu().numpy()
    for i in range(B):
        p2d, p3d = odec.build_correspondences(mask[i], want_ids[i].astype(np.float64), tab)
        assert counts[i] == len(p2d)
        if len(p2d) == 0:
            continue
        uv = odec.mapping_pixel_position_to_original_position(p2d, bb[i], S).astype(np.float32)
        n = counts[i]
        assert np.array_equal(corr[i, 0:2, :n].T, uv)
        assert np.array_equal(corr[i, 2:5, :n].T.view(np.uint32), p3d.astype(np.float32).view(np.uint32))
    # bf16 logits and a sigmoid mask channel in front of the digits (the network's own layout): same ids where the
    # bf16 rounding keeps the arg-max unambiguous -- here simply: runs, and agrees with the fp32 path on its own bf16 input
    lgm = torch.cat([torch.from_numpy(np.where(mask, 2.0, -2.0).astype(np.float32)).cuda().unsqueeze(1), lg], 1)
    c2, n2, k2 = eng.decode_ce(lgm, bb, base=base, n_digits=L, obj_default=2, return_codes=True)
    assert torch.equal(k2.cpu(), codes.cpu()) and torch.equal(n2.cpu(), torch.from_numpy(counts))
    kb = eng.decode_ce(lgm.to(torch.bfloat16), bb, base=base, n_digits=L, obj_default=2, return_codes=True)[2]
    kf = eng.decode_ce(lgm.to(torch.bfloat16).float(), bb, base=base, n_digits=L, obj_default=2, return_codes=True)[2]
    assert torch.equal(kb, kf)


def test_ce_argument_checks(eng):
    import zebrapose_b200 as zp
    tab, _, _ = synth.make_dict(16, seed=9, radius=40.0)
    eng.upload_dict(2, tab, n_bits=16, ignore_bit=0)
    lg = torch.zeros(1, 41, 16, 16, device="cuda")
    with pytest.raises(zp.ZpError):
        eng.decode_ce(lg, np.zeros((1, 4)), base=4, n_digits=10, obj_default=2)      # 4^10 > 65536 classes
    with pytest.raises(ValueError):
        eng.decode_ce(lg, np.zeros((1, 4)), base=8, n_digits=6, obj_default=2)       # 1 + 48 channels > 41
