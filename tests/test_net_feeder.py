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


@pytest.mark.parametrize("H,W,d", [(32, 32, 18), (32, 32, 12), (32, 32, 6), (24, 40, 18), (16, 16, 18), (33, 20, 7)])
def test_tap_conv_equals_dilated_conv(H, W, d):
    """the nine-tap form of a dilated 3x3 convolution (used for the layer cuDNN runs pathologically) is the same sum"""
    g = torch.Generator().manual_seed(H * 100 + d)
    conv = torch.nn.Conv2d(12, 7, 3, 1, d, d)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g))
        conv.bias.copy_(torch.randn(7, generator=g))
        for x in (torch.randn(3, 12, H, W, generator=g),
                  torch.randn(3, 12, H, W, generator=g).contiguous(memory_format=torch.channels_last)):
            want = torch.nn.functional.conv2d(x.double(), conv.weight.double(), conv.bias.double(), 1, d, d)
            got = znet.tap_conv3x3(conv, x)
            assert got.shape == want.shape
            assert float((got.double() - want).abs().max()) < 1e-5 * float(want.abs().max())


def test_tap_form_of_the_network_equals_the_plain_one(run):
    net, img, x, x_128, _ = run
    tapped = znet.build(seed=0, tap_dilation=18)
    assert tapped.aspp.tap_dilation == 18 and net.aspp.tap_dilation == 0
    with torch.no_grad():
        tx, _ = tapped(img[:1])
    assert float((tx - x[:1]).abs().max()) < 1e-5 * max(1.0, float(x[:1].abs().max()))
