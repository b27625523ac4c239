"""Drop-in for the reference module `metric` (zebrapose/metric.py:8-18): ADD / ADI of one pose pair, computed by the
device kernels behind zp_pose_errors (the reference goes through bop_toolkit_lib.pose_error.add / adi, vendored copy
lib/pysixd/pose_error.py:297-336).  The batched form is Engine.pose_errors(); these per-call wrappers upload the
vertices once per distinct array and exist so test.py:465-483 can switch without edits."""
import numpy as np

from .engine import default_engine

_MODEL_SLOT = 255
_model_cache = {"key": None, "ref": None}


def _engine_with_model(vertices):
    eng = default_engine()
    v = np.asarray(vertices)
    key = (id(vertices), v.shape)
    if _model_cache["key"] != key:
        eng.upload_model(_MODEL_SLOT, v)
        _model_cache["key"] = key
        _model_cache["ref"] = vertices          # keep the object alive so id() stays unique
    return eng


def _pose12(R, t):
    return np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(t, np.float64).reshape(3)]).reshape(1, 12)


def Calculate_ADD_Error_BOP(R_GT, t_GT, R_predict, t_predict, vertices):
    """metric.py:8-12 -> pose_error.add(R_predict, t_predict, R_GT, t_GT, vertices)"""
    eng = _engine_with_model(vertices)
    add, _ = eng.pose_errors(_pose12(R_predict, t_predict), _pose12(R_GT, t_GT), obj_default=_MODEL_SLOT, adi=False)
    return float(add.item())


def Calculate_ADI_Error_BOP(R_GT, t_GT, R_predict, t_predict, vertices):
    """metric.py:14-18 -> pose_error.adi(R_predict, t_predict, R_GT, t_GT, vertices)"""
    eng = _engine_with_model(vertices)
    _, adi = eng.pose_errors(_pose12(R_predict, t_predict), _pose12(R_GT, t_GT), obj_default=_MODEL_SLOT, add=False)
    return float(adi.item())
