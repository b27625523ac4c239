"""Pose error metrics for the parity harness (test infrastructure).
ADD / ADI restate lib/pysixd/pose_error.py:297-336 of the reference (numpy + scipy cKDTree)."""
import numpy as np
from scipy.spatial import cKDTree


def transform_pts_Rt(pts, R, t):
    return (R @ pts.T + np.asarray(t).reshape(3, 1)).T


def add(R_est, t_est, R_gt, t_gt, pts):
    """pose_error.py:297-312"""
    return np.linalg.norm(transform_pts_Rt(pts, R_est, t_est) - transform_pts_Rt(pts, R_gt, t_gt), axis=1).mean()


def adi(R_est, t_est, R_gt, t_gt, pts):
    """pose_error.py:315-336"""
    pe = transform_pts_Rt(pts, R_est, t_est)
    pg = transform_pts_Rt(pts, R_gt, t_gt)
    d, _ = cKDTree(pe).query(pg, k=1)
    return d.mean()


def rot_err_deg(R1, R2):
    c = (np.trace(np.asarray(R1).T @ np.asarray(R2)) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))


def trans_err(t1, t2):
    return float(np.linalg.norm(np.asarray(t1).reshape(3) - np.asarray(t2).reshape(3)))
