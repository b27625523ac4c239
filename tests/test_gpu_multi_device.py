"""GPU, needs two devices in ONE process: engines on different devices must not share per-process state (function
attributes are per device, streams belong to the engine's device, the drop-in's dictionary cache is per device)."""
import numpy as np
import pytest
import torch

from workloads import synth

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process(tables):
    import zebrapose_b200 as zp
    tab, nrm = tables["full"]
    crops = [synth.make_crop(tab, nrm, 9300 + i) for i in range(3)]
    logits = np.stack([synth.crop_to_logits(c) for c in crops])
    bboxes = np.stack([c["bbox"] for c in crops])
    K = crops[0]["K"]
    outs = []
    torch.cuda.set_device(0)                       # torch's current device stays 0 while engine 1 works on device 1
    for dev in (1, 0, 1):
        eng = zp.Engine(dev)
        eng.upload_dict(0, tab)
        lg = torch.from_numpy(logits).to("cuda:%d" % dev)
        # bf16 stream decode needs 80 KB of dynamic shared memory, the minimal solver 60 KB: both are per-device attributes
        corr16, counts16 = eng.decode(lg.to(torch.bfloat16), bboxes)
        poses, ninl, status = eng.decode_and_pose_batch(lg, bboxes, K)
        eng.set_solver("fast")
        p2, _, s2 = eng.decode_and_pose_batch(lg, bboxes, K)
        torch.cuda.synchronize(dev)
        assert (status.cpu().numpy() == 0).all() and (s2.cpu().numpy() == 0).all()
        outs.append((poses.cpu().numpy(), ninl.cpu().numpy(), counts16.cpu().numpy()))
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(outs[0], o))
    # the per-crop drop-in follows torch's current device
    from zebrapose_b200.binary_code_helper.CNN_output_to_pose import CNN_outputs_to_object_pose
    from oracle import decode
    c = crops[0]
    mask = decode.threshold_logits(logits[0, 0]).astype(np.uint8)
    code = decode.threshold_logits(logits[0, 1:]).transpose(1, 2, 0)
    d = {float(i): tab[i] for i in range(len(tab))}
    res = []
    for dev in (0, 1, 0):
        torch.cuda.set_device(dev)
        res.append(CNN_outputs_to_object_pose(mask, code, c["bbox"], 128, 2, d, intrinsic_matrix=K))
    torch.cuda.set_device(0)
    assert all(r[2] for r in res) and all(np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) for r in res)
