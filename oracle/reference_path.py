"""The reference's CPU path for one crop, end to end, as test.py:250-273 + CNN_output_to_pose.py:100-160 run it
(test infrastructure / timed CPU baseline).  Two forms with identical results (tests/test_oracle_vs_reference_inplace.py):
`kind() == "reference"` -- the reference's OWN functions imported in place from /root/reference (oracle/ref_inplace.py; only
where that tree exists, i.e. the build container: nothing is copied, the reference sources cannot travel to the GPU box);
`"port"` -- the restatement below, which is what runs on the GPU box.

Step by step: sigmoid + 0.5 threshold of ALL logits into float64 {0,1} arrays (common_ops.py:5-19), NCHW->NHWC
transposes and uint8 mask (test.py:254-257), code -> class id in float64 (class_id_encoder_decoder.py:17-28),
mask.nonzero(), the per-pixel PYTHON LOOP over a dict with two np.array + isnan().any() per pixel
(CNN_output_to_pose.py:53-64; 80 % of the reference's wall time), pixel remap (:34-50), float32 casts,
cv2.solvePnPRansac(EPNP, 150 iterations, 2 px) + cv2.Rodrigues (:155-158).
"""
import numpy as np
import cv2
import torch

from . import decode, ref_inplace


def kind():
    """ "reference" where /root/reference exists (its own functions are imported in place), else "port" """
    return "reference" if ref_inplace.modules() else "port"


def reference_pose_from_logits_ref(logits, bbox, K, dict_float_keys, ignore_bit=0, S=None):
    """the same call sequence as test.py:250-273 through the reference's own functions (imported in place)"""
    co, cnn, _ = ref_inplace.modules()
    lt = torch.from_numpy(np.ascontiguousarray(logits))[None]
    nb = lt.shape[1] - 1
    pred_masks = co.from_output_to_class_mask(lt[:, :1])
    pred_code_images = co.from_output_to_class_binary_code(lt[:, 1:], "BCE", divided_num_each_interation=2, binary_code_length=nb)
    pred_code_images = pred_code_images.transpose(0, 2, 3, 1)
    pred_masks = pred_masks.transpose(0, 2, 3, 1)
    pred_masks = pred_masks.squeeze(axis=-1).astype('uint8')
    code = pred_code_images[0][:, :, :-ignore_bit] if ignore_bit else pred_code_images[0]
    return cnn.CNN_outputs_to_object_pose(pred_masks[0], code, bbox, S or logits.shape[-1], 2, dict_float_keys, intrinsic_matrix=K)


def reference_pose_from_logits(logits, bbox, K, dict_float_keys, ignore_bit=0, S=None):
    """logits float32 [1+n_bits, S, S] (host).  Returns (R, t, success) like CNN_outputs_to_object_pose."""
    lt = torch.from_numpy(np.ascontiguousarray(logits))[None]
    # common_ops.py:5-19 -- sigmoid on the tensor, comparison on the host, float64 outputs
    pm = torch.sigmoid(lt[:, :1]).detach().cpu().numpy()
    pred_mask = np.zeros(pm.shape)
    pred_mask[pm > 0.5] = 1.0
    pc = torch.sigmoid(lt[:, 1:]).detach().cpu().numpy()
    pred_code = np.zeros(pc.shape)
    pred_code[pc > 0.5] = 1.0
    # test.py:254-257
    pred_code = pred_code.transpose(0, 2, 3, 1)
    pred_mask = pred_mask.transpose(0, 2, 3, 1).squeeze(axis=-1).astype("uint8")
    code = pred_code[0][:, :, :-ignore_bit] if ignore_bit else pred_code[0]
    S = S or logits.shape[-1]
    # CNN_output_to_pose.py:110-129
    ids = decode.class_code_images_to_class_id_image(code, 2)
    if pred_mask[0].nonzero()[0].size == 0:
        return [], [], False
    p2d, p3d = decode.build_correspondences_faithful(pred_mask[0], ids, dict_float_keys)
    o2d = decode.mapping_pixel_position_to_original_position(p2d, bbox, S)
    if len(o2d) < 6:
        return [], [], False
    _, rvec, tvec, _ = cv2.solvePnPRansac(p3d.astype(np.float32), o2d.astype(np.float32), np.ascontiguousarray(K),
                                          distCoeffs=None, reprojectionError=2, iterationsCount=150,
                                          flags=cv2.SOLVEPNP_EPNP)
    R, _ = cv2.Rodrigues(rvec)
    return R, tvec, True


# ---- process-pool driver (one worker per host core, cv2.setNumThreads(1)) ----------
_G = {}


def _init(logits, bboxes, Ks, obj, dicts):
    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    _G.update(logits=logits, bboxes=bboxes, Ks=Ks, obj=obj, dicts=dicts, ref=kind() == "reference")


def _work(i):
    fn = reference_pose_from_logits_ref if _G.get("ref") else reference_pose_from_logits
    R, t, ok = fn(_G["logits"][i], _G["bboxes"][i], _G["Ks"][i], _G["dicts"][_G["obj"][i]])
    if not ok:
        return np.zeros(12)
    return np.concatenate([np.asarray(R).ravel(), np.asarray(t).ravel()])


class ReferencePool:
    """fork pool over the crops of a batch; data is inherited by fork, only indices and 12 floats cross the pipe"""

    def __init__(self, logits, bboxes, Ks, obj, tables, cores):
        import multiprocessing as mp
        dicts = [decode.table_to_dict(t, float_keys=True) for t in tables]
        self.n = len(logits)
        self.cores = cores
        _init(logits, bboxes, Ks, obj, dicts)
        self.pool = mp.get_context("fork").Pool(cores, initializer=_init, initargs=(logits, bboxes, Ks, obj, dicts)) if cores > 1 else None

    def run(self, idx):
        if self.pool is None:
            return np.stack([_work(i) for i in idx])
        return np.stack(self.pool.map(_work, list(idx), chunksize=max(1, min(4, len(idx) // self.cores or 1))))

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
