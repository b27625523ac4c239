"""GPU: decode kernel (through the C ABI) vs the golden outputs of the reference and vs the oracle -- bit-exact."""
import numpy as np
import pytest
import torch

from oracle import decode
from workloads import synth
from helpers import GOLDEN_CROPS, regen_crop, as_set

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import zebrapose_b200 as zp
    return zp.Engine(0)


def _oracle_decode(logits, bbox, S, tab, k=0, ext_mask=None):
    mask = decode.threshold_logits(logits[0]).astype(np.uint8) if ext_mask is None else ext_mask
    code = decode.threshold_logits(logits[1:]).transpose(1, 2, 0)
    t = decode.generate_new_corres_table(tab, 16, 16 - k) if k else tab
    if k:
        code = code[:, :, :-k]
    return decode.decode_crop(mask, code, bbox, S, t)


@pytest.mark.parametrize("tag", list(GOLDEN_CROPS))
def test_golden_crops(eng, golden, tables, tag):
    tab, c, logits, S, k = regen_crop(tables, tag)
    eng.upload_dict(1, tab, n_bits=16, ignore_bit=k, nonexist="zero")
    lg = torch.from_numpy(logits)[None].cuda()
    corr, counts, codes = eng.decode(lg, c["bbox"][None], obj_default=1, n_bits=16, ignore_bit=k, return_codes=True)
    n = int(counts.item())
    exp_uv = golden[tag + "_uv"].astype(np.float32)
    exp_xyz = golden[tag + "_xyz"]
    assert n == len(exp_uv)
    assert np.array_equal(codes[0].cpu().numpy(), golden[tag + "_ids"])
    got = corr[0].cpu().numpy()
    # ordered equality (row-major, as mask.nonzero()) and the order-independent multiset the north_star asks for
    assert np.array_equal(got[0:2, :n].T, exp_uv)
    assert np.array_equal(got[2:5, :n].T.view(np.uint32), exp_xyz.view(np.uint32))
    assert np.array_equal(as_set(got[0:2, :n].T, got[2:5, :n].T), as_set(exp_uv, exp_xyz))


def test_table_build_matches_reference_newdict(eng, golden, tables):
    tab = tables["nan20"][0]
    for k in (1, 3, 8):
        eng.upload_dict(2, tab, n_bits=16, ignore_bit=k, nonexist="zero")
        pts, remap = eng.download_tables(2)
        ref = golden["newdict_k%d" % k]
        nanrow = np.isnan(ref).any(1)
        exp = np.where(nanrow[:, None], 0.0, ref).astype(np.float32)
        assert np.array_equal(pts[:, :3].view(np.uint32), exp.view(np.uint32))
        assert np.array_equal(pts[:, 3] == 0, nanrow)
        assert np.array_equal(remap, np.arange(len(ref), dtype=np.uint16))


def test_views_bf16_extmask_multiobj(eng, tables):
    """17-channel tensor consumed through the (mask, code) views the net returns; bf16 logits; external mask;
    several dictionaries in one batch; 18-channel v2 layout (mask, entire mask, bits)."""
    S, B = 128, 6
    tabs = [tables["full"], tables["nan20"]]
    for j, (t, _) in enumerate(tabs):
        eng.upload_dict(10 + j, t, n_bits=16, ignore_bit=0)
    rng = np.random.default_rng(0)
    obj = rng.integers(0, 2, B)
    crops = [synth.make_crop(tabs[obj[i]][0], tabs[obj[i]][1], 9000 + i, S=S) for i in range(B)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    bboxes = np.stack([c["bbox"] for c in crops])
    oid = (10 + obj).astype(np.int32)
    full = torch.from_numpy(logits).cuda()
    mview, cview = torch.split(full, [1, 16], 1)
    assert not cview.is_contiguous()
    corr_a, cnt_a = eng.decode((mview, cview), bboxes, oid)
    corr_b, cnt_b = eng.decode(full.to(torch.bfloat16), bboxes, oid)
    ext = np.stack([np.roll(c["mask"], 3, 1) for c in crops])
    corr_c, cnt_c = eng.decode(full, bboxes, oid, ext_mask=ext)
    v2 = torch.cat([full[:, :1], torch.zeros_like(full[:, :1]), full[:, 1:]], 1)
    corr_d, cnt_d = eng.decode(v2, bboxes, oid, mask_ch=0, bit0_ch=2)
    for i, c in enumerate(crops):
        uv, xyz, _ = _oracle_decode(logits[i], c["bbox"], S, tabs[obj[i]][0])
        for corr, cnt in ((corr_a, cnt_a), (corr_b, cnt_b), (corr_d, cnt_d)):
            n = int(cnt[i])
            g = corr[i].cpu().numpy()
            assert n == len(uv)
            assert np.array_equal(g[0:2, :n].T, uv) and np.array_equal(g[2:5, :n].T.view(np.uint32), xyz.view(np.uint32))
        uv, xyz, _ = _oracle_decode(logits[i], c["bbox"], S, tabs[obj[i]][0], ext_mask=ext[i])
        n = int(cnt_c[i])
        g = corr_c[i].cpu().numpy()
        assert n == len(uv) and np.array_equal(g[0:2, :n].T, uv) and np.array_equal(g[2:5, :n].T.view(np.uint32), xyz.view(np.uint32))


@pytest.mark.parametrize("S,layout", [(128, "channels_last"), (100, "contig"), (50, "contig"), (256, "contig"), (32, "contig")])
def test_generic_and_odd_sizes(eng, tables, S, layout):
    """generic (scalar, two-kernel) path: channels-last strides, crop sizes the cluster path cannot take"""
    tab, nrm = tables["nan20"]
    eng.upload_dict(3, tab, n_bits=16, ignore_bit=2)
    B = 3
    crops = [synth.make_crop(tab, nrm, 700 + i, S=S) for i in range(B)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    bboxes = np.stack([c["bbox"] for c in crops]).astype(np.float64)
    bboxes[0] += [0.25, -0.5, 0.75, 0.5]                      # float boxes go through the float64 path
    lg = torch.from_numpy(logits).cuda()
    if layout == "channels_last":
        lg = lg.to(memory_format=torch.channels_last)
        assert lg.stride(3) != 1
    corr, counts, codes = eng.decode(lg, bboxes, obj_default=3, ignore_bit=2, return_codes=True)
    for i, c in enumerate(crops):
        uv, xyz, ids = _oracle_decode(logits[i], bboxes[i], S, tab, k=2)
        n = int(counts[i])
        g = corr[i].cpu().numpy()
        assert n == len(uv)
        assert np.array_equal(codes[i].cpu().numpy().astype(np.int64), ids)
        assert np.array_equal(g[0:2, :n].T, uv) and np.array_equal(g[2:5, :n].T.view(np.uint32), xyz.view(np.uint32))


def test_edge_masks_and_cap(eng, tables):
    tab, nrm = tables["full"]
    eng.upload_dict(4, tab)
    S = 128
    c = synth.make_crop(tab, nrm, 31, S=S)
    logits = np.stack([synth.crop_to_logits(c)] * 3)
    logits[0, 0] = -3.0                                       # empty mask
    logits[1, 0] = 3.0                                        # full mask: every pixel, maximum size
    logits[2, 0] = -3.0
    logits[2, 0, 5, 7:12] = 2.0                               # five pixels
    lg = torch.from_numpy(logits).cuda()
    bb = np.stack([c["bbox"], [-5, -5, 100, 100], [0, 0, 0, 0]])
    corr, counts = eng.decode(lg, bb, obj_default=4)
    assert counts.tolist() == [0, S * S, 5]
    uv, xyz, _ = _oracle_decode(logits[1], bb[1], S, tab)
    g = corr[1].cpu().numpy()
    assert np.array_equal(g[0:2].T, uv) and np.array_equal(g[2:5].T.view(np.uint32), xyz.view(np.uint32))
    assert g[0].min() < 0                                     # truncation toward zero on negative coordinates
    assert np.array_equal(corr[2, 0:2, :5].cpu().numpy(), np.zeros((2, 5), np.float32))   # zero box -> (0,0)
    # a cap smaller than the list clips the list, the count still reports every masked pixel
    corr2, counts2 = eng.decode(lg, bb, obj_default=4, cap=1024)
    assert counts2.tolist() == [0, S * S, 5]
    assert np.array_equal(corr2[1].cpu().numpy(), g[:, :1024])
    # NaN / zero logits are background
    lg2 = lg.clone()
    lg2[1, 0, :4] = float("nan")
    lg2[1, 0, 4:8] = 0.0
    _, counts3 = eng.decode(lg2, bb, obj_default=4)
    assert counts3[1].item() == S * S - 8 * S


def test_hamming_mode_against_spec(eng, tables):
    """north_star extension (parity unpinned by the reference): Hamming-nearest remap of non-existing codes"""
    tab, nrm = tables["nan20"]
    for k in (0, 3):
        eng.upload_dict(5, tab, n_bits=16, ignore_bit=k, nonexist="hamming")
        pts, remap = eng.download_tables(5)
        exp_pts, exp_remap, exists = decode.ignore_bit_table_hamming(tab, k)
        assert np.array_equal(remap, exp_remap)
        assert np.array_equal(pts[:, :3].view(np.uint32), exp_pts.astype(np.float32).view(np.uint32))
        assert np.array_equal(pts[:, 3] != 0, exists)
    c = synth.make_crop(tab, nrm, 77)
    logits = synth.crop_to_logits(c)
    corr, counts = eng.decode(torch.from_numpy(logits)[None].cuda(), c["bbox"][None], obj_default=5, ignore_bit=3)
    uv, xyz, _ = decode.decode_crop(decode.threshold_logits(logits[0]).astype(np.uint8),
                                    decode.threshold_logits(logits[1:14]).transpose(1, 2, 0), c["bbox"], 128, exp_pts)
    n = int(counts.item())
    g = corr[0].cpu().numpy()
    assert n == len(uv) and np.array_equal(g[2:5, :n].T.view(np.uint32), xyz.view(np.uint32))


def test_full_size_properties(eng, tables):
    """BASELINE config 2 size (64 crops): size-independent properties instead of an oracle pass"""
    B, S = 64, 128
    logits, bboxes, Ks, obj, tabs, crops = synth.make_batch(B, S=S, n_dicts=3, seed=1002, K=synth.YCBV_K)
    for j, t in enumerate(tabs):
        eng.upload_dict(20 + j, t)
    lg = torch.from_numpy(logits).cuda()
    corr, counts, codes = eng.decode(lg, bboxes, (20 + obj).astype(np.int32), return_codes=True)
    m = (lg[:, 0] > 0)
    assert torch.equal(counts.long(), m.flatten(1).sum(1))                       # count == mask pixels
    w = (2 ** torch.arange(15, -1, -1, device="cuda")).view(1, 16, 1, 1)
    ids = ((lg[:, 1:] > 0) * w).sum(1)
    assert torch.equal(codes.to(torch.int64), ids)                               # every code
    for i in (0, 17, 63):                                                        # lists are row-major sorted
        n = int(counts[i])
        v = corr[i, 1, :n].cpu().numpy()
        assert (np.diff(v) >= 0).all()                                           # rows never go backwards
    corr2, counts2 = eng.decode(lg, bboxes, (20 + obj).astype(np.int32))         # idempotent / deterministic
    assert torch.equal(counts, counts2)
    for i in range(B):
        n = int(counts[i])
        assert torch.equal(corr[i, :, :n], corr2[i, :, :n])


@pytest.mark.parametrize("B,S,dtype,ext,k", [(1, 128, "f32", False, 0), (3, 128, "bf16", False, 4), (70, 128, "f32", False, 0),
                                             (5, 64, "f32", True, 0), (2, 100, "f32", False, 2), (40, 128, "bf16", True, 0),
                                             (2, 20, "f32", False, 0), (300, 32, "f32", False, 0)])
def test_decode_paths_identical(eng, tables, B, S, dtype, ext, k):
    """the decode paths (fused streaming kernel with 1/3/8/automatic runs per CTA, single-run fused with a DSMEM cluster
    exchange or independent CTAs, generic strided, fused TMA ring, two-kernel stream + emit) write identical lists,
    counts and codes; crop 0 is also checked against the oracle"""
    tab, nrm = tables["nan20"]
    eng.upload_dict(6, tab, n_bits=16, ignore_bit=k)
    crops = [synth.make_crop(tab, nrm, 5000 + (i % 7), S=S) for i in range(B)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    rng = np.random.default_rng(B * 1000 + S)
    logits[:, 0] *= np.where(rng.random((B, 1, 1)) < 0.15, -1.0, 1.0)          # some inverted (large / small) masks
    if B > 2:
        logits[1, 0] = -1.0                                                   # an empty crop in the middle
    bboxes = np.stack([c["bbox"] for c in crops])
    lg = torch.from_numpy(logits).cuda()
    if dtype == "bf16":
        lg = lg.to(torch.bfloat16)
    em = (rng.random((B, S, S)) < 0.4).astype(np.uint8) if ext else None
    outs = []
    try:
        for path in (0, 1, 2, 3, 4, 6, 101, 103, 108, 100):
            eng.set_decode_path(path)
            corr, counts, codes = eng.decode(lg, bboxes, obj_default=6, ignore_bit=k, ext_mask=em, return_codes=True)
            outs.append((corr.cpu().numpy(), counts.cpu().numpy(), codes.cpu().numpy()))
    finally:
        eng.set_decode_path(0)
    c0, n0, k0 = outs[0]
    for c1, n1, k1 in outs[1:]:
        assert np.array_equal(n0, n1) and np.array_equal(k0, k1)
        for i in range(B):
            assert np.array_equal(c0[i, :, :n0[i]].view(np.uint32), c1[i, :, :n1[i]].view(np.uint32))
    lg0 = lg[0].float().cpu().numpy()
    uv, xyz, ids = _oracle_decode(lg0, bboxes[0], S, tab, k=k, ext_mask=None if em is None else em[0])
    assert n0[0] == len(uv) and np.array_equal(k0[0].astype(np.int64), ids)
    assert np.array_equal(c0[0, 0:2, :n0[0]].T, uv) and np.array_equal(c0[0, 2:5, :n0[0]].T.view(np.uint32), xyz.view(np.uint32))


def test_bad_object_ids_are_memory_safe(tables):
    """obj ids outside [0,256), slots that were never uploaded and slots holding a SHORTER table than the call's code
    length must not read out of bounds: they decode against "no code exists" / zero-padded rows ((0,0,0) points) and the
    crops with a good id are unaffected"""
    import zebrapose_b200 as zp
    eng = zp.Engine(0)
    tab, nrm = tables["full"]
    eng.upload_dict(0, tab)
    eng.upload_dict(7, tab, ignore_bit=6)                 # 1024-row table in slot 7
    crops = [synth.make_crop(tab, nrm, 9100 + i) for i in range(5)]
    logits = torch.from_numpy(np.stack([synth.crop_to_logits(c) for c in crops])).cuda()
    bboxes = np.stack([c["bbox"] for c in crops])
    ids = torch.tensor([0, -3, 100000, 5, 7], dtype=torch.int32).cuda()     # ok | negative | huge | empty slot | short table
    good_corr, good_counts = eng.decode(logits, bboxes)
    corr, counts = eng.decode(logits, bboxes, ids)
    torch.cuda.synchronize()
    assert torch.equal(counts, good_counts)
    n = int(counts[0])
    assert torch.equal(corr[0, :, :n], good_corr[0, :, :n])
    for b in (1, 2, 3):
        nb = int(counts[b])
        assert torch.equal(corr[b, 0:2, :nb], good_corr[b, 0:2, :nb]) and float(corr[b, 2:5, :nb].abs().max()) == 0.0
    assert torch.isfinite(corr[4, :, :int(counts[4])]).all()
    r = eng.ransac(corr, counts, crops[0]["K"])
    st = r["status"].cpu().numpy()
    assert st[0] == 0 and (st[1:4] == 3).all()            # no model from all-zero 3D points
