"""Where the time of the configs[4] network body goes (torch / cuDNN, not this repo's kernels): per-stage CUDA-event
times of workloads/net.py at B crops, bf16, for cudnn.benchmark off / on and channels_last / NCHW.

    python tools/prof_net_body.py [B=128]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from workloads import net as znet

GFLOP = {"stem": 0.0, "layer1": 0.0, "layer2": 0.0, "layer4": 0.0, "layer5": 0.0, "aspp": 0.0, "up1": 0.0, "up2": 0.0}


def stages(net, img):
    r, a = net.resnet, net.aspp
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    with torch.no_grad():
        ev[0].record()
        x_128 = r.resnet_layer_1(img); ev[1].record()
        x_64 = r.resnet_layer_2(x_128); ev[2].record()
        x_32 = r.resnet_layer_3(x_64); ev[3].record()
        x_16 = r.layer4(x_32); ev[4].record()
        feat = r.layer5(x_16); ev[5].record()
        if a.aspp_nchw:
            feat = feat.contiguous()
        br = [znet._cb(a.conv_1x1_1, a.bn_conv_1x1_1, feat)]
        for i in (1, 2, 3):
            conv, bn = getattr(a, "conv_3x3_%d" % i), getattr(a, "bn_conv_3x3_%d" % i)
            if 0 < a.tap_dilation <= conv.dilation[0]:
                br.append(torch.relu_(bn(znet.tap_conv3x3(conv, feat))))
            else:
                br.append(znet._cb(conv, bn, feat))
        pooled = znet._cb(a.conv_1x1_2, a.bn_conv_1x1_2, feat.mean((2, 3), keepdim=True))
        br.append(pooled.expand(-1, -1, feat.shape[2], feat.shape[3]))
        y = znet._cb(a.conv_1x1_3, a.bn_conv_1x1_3, torch.cat(br, 1)); ev[6].record()
        u1 = a.upsample_1(y); ev[7].record()
        a.upsample_2(torch.cat([u1, x_64], 1)); ev[8].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(8)]


def aspp_ops(net, img):
    """each ASPP operation alone (CUDA events, third of three runs)"""
    r, a = net.resnet, net.aspp
    out = {}
    with torch.no_grad():
        feat = r.layer5(r.layer4(r.resnet_layer_3(r.resnet_layer_2(r.resnet_layer_1(img)))))
        if a.aspp_nchw:
            feat = feat.contiguous()
        ops = {"conv_1x1_1": lambda: a.conv_1x1_1(feat), "conv_3x3_d6": lambda: a.conv_3x3_1(feat),
               "conv_3x3_d12": lambda: a.conv_3x3_2(feat), "conv_3x3_d18": lambda: a.conv_3x3_3(feat),
               "taps_d6": lambda: znet.tap_conv3x3(a.conv_3x3_1, feat), "taps_d12": lambda: znet.tap_conv3x3(a.conv_3x3_2, feat),
               "taps_d18": lambda: znet.tap_conv3x3(a.conv_3x3_3, feat),
               "mean": lambda: feat.mean((2, 3), keepdim=True)}
        cat = torch.cat([a.conv_1x1_1(feat)] * 5, 1)
        ops["conv_1x1_3"] = lambda: a.conv_1x1_3(cat)
        for k, fn in ops.items():
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            out[k] = round(e0.elapsed_time(e1), 3)
    return out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    names = list(GFLOP)
    for bench, cl, aspp_nchw, taps in ((False, True, False, 0), (False, False, False, 0), (True, True, False, 0),
                                       (False, True, True, 0), (False, True, False, 18), (False, True, False, 6)):
        if True:
            torch.backends.cudnn.benchmark = bench
            net = znet.build(seed=0, device="cuda", dtype=torch.bfloat16, fold=True, aspp_nchw=aspp_nchw, tap_dilation=taps)
            img = znet.images(B, seed=0, device="cuda", dtype=torch.bfloat16)
            if not cl:
                net = net.to(memory_format=torch.contiguous_format)
                img = img.contiguous()
            for _ in range(2):
                stages(net, img)
            t = [0.0] * 8
            for _ in range(3):
                t = [a + b / 3 for a, b in zip(t, stages(net, img))]
            print(json.dumps({"B": B, "cudnn_benchmark": bench, "channels_last": cl, "aspp_nchw": aspp_nchw, "tap_dilation": taps, "ms_total": round(sum(t), 2),
                              "tflops": round(109.1 * B / sum(t), 1),
                              "ms": {n: round(v, 2) for n, v in zip(names, t)}}), flush=True)
            if cl and not bench and not taps:
                print(json.dumps({"aspp_ops_ms": aspp_ops(net, img), "aspp_nchw": aspp_nchw}), flush=True)
            del net, img
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
