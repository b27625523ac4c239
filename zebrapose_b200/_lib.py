"""ctypes binding of libzebrapose_b200.so (include/zebrapose_b200.h).  No CPU fallback: a missing library or a missing
CUDA device raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libzebrapose_b200.so")

DTYPE_F32, DTYPE_BF16 = 0, 1
NONEXIST = {"zero": 0, "hamming": 1}
SAMPLER = {"cv2": 0, "philox": 1}
SELECT = {"cv2_replay": 0, "argmax": 1}
FINAL = {"epnp": 0, "epnp+gn": 1}
SOLVER = {"cv2": 0, "fast": 1}
RESIZE = {"crop_resize": 0, "crop_square_resize": 1, "crop_resize_by_warp_affine": 2, "none": 3}
STATUS_OK, STATUS_NO_MASK, STATUS_TOO_FEW, STATUS_NO_MODEL = 0, 1, 2, 3

_vp, _i, _i64, _u64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double

# name -> (restype, argtypes); mirrors include/zebrapose_b200.h one to one
SIGNATURES = {
    "zp_version": (_i, []),
    "zp_create": (_i, [C.POINTER(_vp), _i]),
    "zp_destroy": (None, [_vp]),
    "zp_last_error": (C.c_char_p, [_vp]),
    "zp_upload_tables": (_i, [_vp, _i, _vp, _i, _i, _i]),
    "zp_download_tables": (_i, [_vp, _i, _vp, _vp]),
    "zp_decode": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_i64), _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "zp_decode_ce": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_i64), _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "zp_make_samples": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _u64, _vp, _vp]),
    "zp_solve_minimal": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "zp_score": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _f, _vp, _vp]),
    "zp_ransac": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _f, _d, _i, _u64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                       _vp]),
    "zp_set_solver": (_i, [_vp, _i]),
    "zp_set_waves": (_i, [_vp, _i, _vp]),
    "zp_set_final_form": (_i, [_vp, _i]),
    "zp_set_exact_ties": (_i, [_vp, _i]),
    "zp_pose_batch_device": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_i64), _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _d, _i,
                                  _u64, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "zp_pose_batch_host": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _f, _d, _i, _u64, _i, _i,
                                _vp, _vp, _vp]),
    "zp_pose_batch_host_async": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _f, _d, _i, _u64, _i, _i,
                                      _vp, _vp, _vp]),
    "zp_sync": (_i, [_vp]),
    "zp_remap_pixels": (_i, [_vp, _vp, _i64, _vp, _i, _vp, _vp]),
    "zp_codes_to_ids": (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp]),
    "zp_launch_count": (_i64, [_vp]),
    "zp_set_kernel_timing": (_i, [_vp, _i]),
    "zp_kernel_time": (_i, [_vp, C.c_char_p, C.POINTER(_d), C.POINTER(_i64)]),
    "zp_set_decode_path": (_i, [_vp, _i]),
    "zp_debug_buffer": (_i, [_vp, _vp]),
    "zp_set_score_groups": (_i, [_vp, _i, _i]),
    "zp_debug_clocks": (_i, [_vp, _vp]),
    "zp_fp32_peak_probe": (_i, [_vp, _i, C.POINTER(_d)]),
    "zp_fp32x2_peak_probe": (_i, [_vp, _i, C.POINTER(_d)]),
    "zp_fp64_peak_probe": (_i, [_vp, _i, C.POINTER(_d)]),
    "zp_final_bbox": (_i, [_vp, _vp, _i, _d, _i, _d, _d, _vp, _vp]),
    "zp_crop_input": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "zp_upload_model": (_i, [_vp, _i, _vp, _i]),
    "zp_pose_errors": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "zp_upload_head": (_i, [_vp, _vp, _vp, _i, _i]),
    "zp_head_decode": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
}

_lib = None


class ZpError(RuntimeError):
    pass


def load():
    """dlopen the library (building nothing: `python -m zebrapose_b200._build` or __graft_entry__.build() does that)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZpError("libzebrapose_b200.so not built (%s missing): run `python zebrapose_b200/_build.py`; "
                      "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Context:
    """One zp_ctx (per device)."""

    def __init__(self, device=0):
        self.lib = load()
        self.handle = _vp()
        rc = self.lib.zp_create(C.byref(self.handle), int(device))
        if rc != 0:
            msg = self.lib.zp_last_error(None)
            raise ZpError("zp_create(device=%d) failed: %s" % (device, msg.decode() if msg else rc))
        self.device = int(device)

    def check(self, rc, what):
        if rc != 0:
            msg = self.lib.zp_last_error(self.handle)
            raise ZpError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.zp_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
