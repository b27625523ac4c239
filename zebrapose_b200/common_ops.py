"""Mirror of zebrapose/common_ops.py:5-30.  The reference pulls all logits to the host and thresholds there; here the
comparison runs on the device the logits already live on and only the {0,1} result is copied.
BCE / L1 heads: for the reference's t = 0.5, sigmoid(x) > t is float32(x) > 0; any other t evaluates the
reference's own expression torch.sigmoid(x) > t on the device (t <= 0 -> all ones, t >= 1 -> all zeros).
CE heads (ablation configs, class_base = divided_num_each_interation): softmax over each group of `base` consecutive
channels, argmax (first maximum on ties) -- the same torch softmax the reference applies, on the tensor's device, then
argmax there instead of on the host.
The batched path (Engine.decode_and_pose_batch) never calls these: its decode kernel thresholds in registers."""
import numpy as np
import torch


def _threshold(pred, thershold):
    x = pred.detach()
    if thershold == 0.5:                  # sigmoid(x) > 0.5  <=>  float32(x) > 0 (differs only for 0 < x < 8.94e-8)
        return (x.to(torch.float32) > 0).to(torch.float64).cpu().numpy()
    # any other threshold (incl. <= 0 and >= 1): exactly the reference's expression, evaluated on the tensor's device
    return (torch.sigmoid(x) > thershold).to(torch.float64).cpu().numpy()


def from_output_to_class_mask(pred_mask_prob, thershold=0.5):
    return _threshold(pred_mask_prob, thershold)


def from_output_to_class_binary_code(pred_code_prob, BinaryCode_Loss_Type, thershold=0.5, divided_num_each_interation=2,
                                     binary_code_length=16):
    if BinaryCode_Loss_Type in ("BCE", "L1"):
        return _threshold(pred_code_prob, thershold)
    if BinaryCode_Loss_Type == "CE":                  # common_ops.py:21-30
        x = pred_code_prob.detach()
        base = int(divided_num_each_interation)
        p = torch.softmax(x.reshape(-1, base, x.shape[2], x.shape[3]), dim=1)
        code = torch.argmax(p, dim=1, keepdim=True)   # first maximal index, as numpy.argmax
        return code.reshape(-1, int(binary_code_length), x.shape[2], x.shape[3]).cpu().numpy()
    raise ValueError("unknown BinaryCode_Loss_Type %r" % (BinaryCode_Loss_Type,))


def get_batch_size(second_dataset_ratio, batch_size):
    second = int(batch_size * second_dataset_ratio)
    return batch_size - second, second
