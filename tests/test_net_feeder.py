"""CPU: workloads/net.py (the network body that feeds BASELINE configs[4]) against the reference model's own forward pass.

tests/golden/golden_net_v1.npz holds logits of the REFERENCE `BinaryCodeNet_Deeplab(34, 16, 2, concat=True,
output_kernel_size=1)` after it loaded (strict) the state dict exported by `workloads.net.build(0)`; here the feeder is
rebuilt from the seed and must reproduce them.  Tolerance: float32 convolutions, 40 layers, logits of std 0.06 ->
1e-5 absolute (observed 3e-8 in the generating run; thread count changes the summation order)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from workloads import net as znet

ATOL = 1e-5
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_net_v1.npz"))


def _digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.digest()


@pytest.fixture(scope="module")
def run():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    net = znet.build(seed=0)
    img = znet.images(2, seed=0)
    with torch.no_grad():
        x, x_128 = net(img)
        logits = net.tail(torch.cat([x, x_128], 1))
    return net, img, x, x_128, logits


def test_seeded_inputs_and_weights_are_the_fixture_ones(run):
    net, img, *_ = run
    assert hashlib.sha256(img.numpy().tobytes()).digest() == G["img_sha256"].tobytes()
    assert _digest(net.state_dict()) == G["weights_sha256"].tobytes()


def test_feeder_reproduces_reference_forward(run):
    _, _, x, x_128, logits = run
    assert x.shape == (2, 256, 128, 128) and x_128.shape == (2, 64, 128, 128)
    lg = logits.numpy()
    np.testing.assert_allclose(lg[:, :, ::8, ::8], G["logits_sub"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(lg[0, :, 77, :], G["logits_row"], rtol=0, atol=ATOL)
    want = np.unpackbits(G["bits_crop0"]).reshape(17, 128, 128).astype(bool)
    sure = np.abs(lg[0]) > ATOL
    assert sure.mean() > 0.999
    assert np.array_equal((lg[0] > 0)[sure], want[sure])


def test_reference_checkpoint_round_trip(run):
    net = run[0]
    sd = net.reference_state_dict()
    # the trunk tensors appear under both of the reference's names (model/resnet.py:184-196)
    assert "net.resnet.resnet.0.weight" in sd and "net.resnet.resnet_layer_1.0.weight" in sd
    assert "net.resnet.resnet.5.3.bn2.running_var" in sd and "net.aspp.conv_1x1_4.bias" in sd
    assert sd["net.resnet.resnet.4.2.conv2.weight"] is sd["net.resnet.resnet_layer_2.1.2.conv2.weight"]
    other = znet.ZebraNetBody().eval()
    other.load_reference_state_dict(sd)
    assert _digest(other.state_dict()) == _digest(net.state_dict())
    bad = dict(sd)
    del bad["net.aspp.upsample_2.0.weight"]
    with pytest.raises(RuntimeError):
        other.load_reference_state_dict(bad)


def test_folded_batchnorm_equals_unfolded(run):
    _, img, x, x_128, _ = run
    folded = znet.build(seed=0, fold=True)
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in folded.modules())
    with torch.no_grad():
        fx, fx_128 = folded(img[:1])
    # folding changes the rounding of every layer; activations are O(0.1 .. 1)
    assert float((fx - x[:1]).abs().max()) < 1e-4 * max(1.0, float(x[:1].abs().max()))
    assert float((fx_128 - x_128[:1]).abs().max()) < 1e-4 * max(1.0, float(x_128[:1].abs().max()))
