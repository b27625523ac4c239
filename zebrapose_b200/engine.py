"""Host side of the B200 pose path: device-resident batched API over the C ABI (include/zebrapose_b200.h).

PyTorch is used for device memory, streams and torch.distributed only; all compute is in libzebrapose_b200.so.
There is no CPU fallback: constructing an Engine without the library or without a CUDA device raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_DT = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16}


def dict_to_table(d, n_bits):
    """Reference dictionary (load_dict_class_id_3D_points / generate_new_corres_dict: keys float or int, values
    (3,) or (1,3) float64, NaN = non-existing) -> float64 [2^n_bits, 3] table.  Missing keys become NaN rows."""
    if isinstance(d, np.ndarray):
        t = np.ascontiguousarray(d, dtype=np.float64)
        if t.shape != (1 << n_bits, 3):
            raise ValueError("table must have shape (2^n_bits, 3), got %s" % (t.shape,))
        return t
    n = 1 << n_bits
    if len(d) == n:
        try:
            return np.ascontiguousarray(np.stack([np.asarray(d[i], np.float64).reshape(3) for i in range(n)]))
        except KeyError:
            pass
    tab = np.full((n, 3), np.nan)
    for k, v in d.items():
        tab[int(k)] = np.asarray(v, np.float64).reshape(3)
    return tab


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One context per device.  All tensor arguments live on that device; every call is asynchronous on the current
    torch CUDA stream unless stated otherwise."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.ZpError("zebrapose_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        self.ctx = _lib.Context(self.device.index)
        self.lib = self.ctx.lib
        self._slots = {}
        self._inflight = []              # host buffers of asynchronous submissions, kept alive until sync()
        self._dev_out = {}               # batch size -> persistent outputs of the graph-replayed entry
        self._graph_refs = {}            # tensors named by captured graphs

    # ------------------------------------------------------------------ dictionaries
    def upload_dict(self, obj_id, table_or_dict, n_bits=16, ignore_bit=0, nonexist="zero"):
        """Replaces the per-crop dict look-ups (CNN_output_to_pose.py:58-62) + generate_new_corres_dict
        (generate_new_dict.py:4-33): uploads the full n_bits dictionary; the ignore-bit parent table and the
        non-existing-code handling are built by the library.  Synchronises."""
        tab = dict_to_table(table_or_dict, n_bits)
        rc = self.lib.zp_upload_tables(self.ctx.handle, int(obj_id), tab.ctypes.data_as(C.c_void_p), int(n_bits),
                                       int(ignore_bit), _lib.NONEXIST[nonexist])
        self.ctx.check(rc, "zp_upload_tables")
        self._slots[int(obj_id)] = (int(n_bits), int(ignore_bit), nonexist)

    def download_tables(self, obj_id):
        n_bits, k, _ = self._slots[int(obj_id)]
        n = 1 << (n_bits - k)
        pts = np.empty((n, 4), np.float32)
        remap = np.empty(n, np.uint16)
        rc = self.lib.zp_download_tables(self.ctx.handle, int(obj_id), pts.ctypes.data_as(C.c_void_p),
                                         remap.ctypes.data_as(C.c_void_p))
        self.ctx.check(rc, "zp_download_tables")
        return pts, remap

    def set_decode_path(self, path):
        """0 auto (fused streaming kernel) | 1 single-run fused, cluster exchange | 2 generic strided | 3 fused TMA ring |
        4 two kernels | 6 single-run fused, independent CTAs | 100+r: r runs per CTA -- identical results"""
        self.ctx.check(self.lib.zp_set_decode_path(self.ctx.handle, int(path)), "zp_set_decode_path")

    def set_solver(self, solver="cv2"):
        """minimal solver of the RANSAC hypotheses: "cv2" = exact replay of OpenCV's EPnP arithmetic (default; hypotheses
        bit-identical to cv2.solvePnP), "fast" = the float64 bisection solver (other null-space basis for 4/5 points)"""
        self.ctx.check(self.lib.zp_set_solver(self.ctx.handle, _lib.SOLVER[solver]), "zp_set_solver")

    def set_waves(self, sizes=None):
        """hypotheses per RANSAC wave (None / empty = automatic); results do not depend on the plan"""
        arr = np.ascontiguousarray(np.asarray(sizes if sizes else [], np.int32))
        self.ctx.check(self.lib.zp_set_waves(self.ctx.handle, int(arr.size), arr.ctypes.data_as(C.c_void_p)), "zp_set_waves")

    def set_exact_ties(self, on=True):
        """near-ties of cv2's record rule decided on exact (cv2-arithmetic) inlier counts (default) or on the FP32 counts;
        on=2 (test aid): exact re-counts without parking the early low-count near-ties"""
        self.ctx.check(self.lib.zp_set_exact_ties(self.ctx.handle, int(on)), "zp_set_exact_ties")

    def set_final_form(self, form=0):
        """final solve split into three kernels (2; the default 0 for final="epnp"), as a 4-CTA cluster per crop (4) or one CTA
        per crop (1): forms 1 and 4 give identical bits, the split form the same pose to rounding"""
        self.ctx.check(self.lib.zp_set_final_form(self.ctx.handle, int(form)), "zp_set_final_form")

    def set_score_groups(self, groups=0, hyp_chunk=0):
        """scheduling knobs of the scoring kernel (include/zebrapose_b200.h); results do not depend on them"""
        self.ctx.check(self.lib.zp_set_score_groups(self.ctx.handle, int(groups), int(hyp_chunk)), "zp_set_score_groups")

    # ------------------------------------------------------------------ decode
    def decode(self, logits, bboxes, obj_ids=None, *, obj_default=0, mask_ch=0, bit0_ch=1, n_bits=16, ignore_bit=0,
               ext_mask=None, return_codes=False, cap=None):
        """logits: cuda tensor [B,C,S,S] (fp32 | bf16, any strides) or the tuple (mask_logits, code_logits) the
        reference network returns (views of one tensor, model/BinaryCodeNet.py:172).  bboxes [B,4] (x,y,w,h).
        Returns corr f32 [B,5,cap], counts i32 [B] (, codes u16 [B,S,S])."""
        if isinstance(logits, (tuple, list)):
            logits, mask_ch, bit0_ch = self._join_views(*logits)
        if logits.dim() != 4 or logits.shape[2] != logits.shape[3]:
            raise ValueError("logits must be [B,C,S,S]")
        if logits.dtype not in _DT:
            raise TypeError("logits dtype %s not supported (float32 | bfloat16)" % logits.dtype)
        if logits.device != self.device:
            raise ValueError("logits live on %s, engine on %s" % (logits.device, self.device))
        B, Cc, S, _ = logits.shape
        if bit0_ch + (n_bits - ignore_bit) > Cc or mask_ch >= Cc:
            raise ValueError("channel layout exceeds the %d channels of logits" % Cc)
        cap = int(cap or ((S * S + 3) // 4) * 4)
        bb = torch.as_tensor(bboxes).to(device=self.device, dtype=torch.float64).contiguous().reshape(B, 4)
        oid = None
        if obj_ids is not None:
            oid = torch.as_tensor(obj_ids).to(device=self.device, dtype=torch.int32).contiguous()
        em = None
        if ext_mask is not None:
            em = torch.as_tensor(ext_mask).to(device=self.device)
            em = (em != 0).to(torch.uint8).contiguous().reshape(B, S, S)
        corr = torch.empty((B, 5, cap), dtype=torch.float32, device=self.device)
        counts = torch.empty((B,), dtype=torch.int32, device=self.device)
        codes = torch.empty((B, S, S), dtype=torch.uint16, device=self.device) if return_codes else None
        strides = (C.c_int64 * 4)(*logits.stride())
        rc = self.lib.zp_decode(self.ctx.handle, _ptr(logits), _DT[logits.dtype], B, S, strides, int(mask_ch),
                                int(bit0_ch), int(n_bits), int(ignore_bit), _ptr(em), _ptr(bb), _ptr(oid),
                                int(obj_default), _ptr(codes), _ptr(corr), cap, _ptr(counts), _stream(self.device))
        self.ctx.check(rc, "zp_decode")
        return (corr, counts, codes) if return_codes else (corr, counts)

    def decode_ce(self, logits, bboxes, obj_ids=None, *, base, n_digits, obj_default=0, mask_ch=0, digit0_ch=1, ext_mask=None,
                  return_codes=False, cap=None):
        """decode() for the CE heads of the ablation configs (common_ops.py:21-30): `n_digits` groups of `base`
        consecutive code channels, digit = first maximum of the group's float32 softmax, id in base `base`."""
        if isinstance(logits, (tuple, list)):
            logits, mask_ch, digit0_ch = self._join_views(*logits)
        if logits.dim() != 4 or logits.shape[2] != logits.shape[3] or logits.dtype not in _DT:
            raise ValueError("logits must be float32 | bfloat16 [B,C,S,S]")
        B, Cc, S, _ = logits.shape
        if digit0_ch + base * n_digits > Cc or mask_ch >= Cc:
            raise ValueError("channel layout exceeds the %d channels of logits" % Cc)
        cap = int(cap or ((S * S + 3) // 4) * 4)
        bb = torch.as_tensor(bboxes).to(device=self.device, dtype=torch.float64).contiguous().reshape(B, 4)
        oid = None if obj_ids is None else torch.as_tensor(obj_ids).to(device=self.device, dtype=torch.int32).contiguous()
        em = None
        if ext_mask is not None:
            em = (torch.as_tensor(ext_mask).to(device=self.device) != 0).to(torch.uint8).contiguous().reshape(B, S, S)
        corr = torch.empty((B, 5, cap), dtype=torch.float32, device=self.device)
        counts = torch.empty((B,), dtype=torch.int32, device=self.device)
        codes = torch.empty((B, S, S), dtype=torch.uint16, device=self.device) if return_codes else None
        strides = (C.c_int64 * 4)(*logits.stride())
        rc = self.lib.zp_decode_ce(self.ctx.handle, _ptr(logits), _DT[logits.dtype], B, S, strides, int(mask_ch),
                                   int(digit0_ch), int(base), int(n_digits), _ptr(em), _ptr(bb), _ptr(oid), int(obj_default),
                                   _ptr(codes), _ptr(corr), cap, _ptr(counts), _stream(self.device))
        self.ctx.check(rc, "zp_decode_ce")
        return (corr, counts, codes) if return_codes else (corr, counts)

    @staticmethod
    def _join_views(mask_logits, code_logits):
        """(mask, code) views of one [B,C,S,S] tensor -> (base tensor view, mask_ch, bit0_ch) without a copy."""
        m, c = mask_logits, code_logits
        same = (m.untyped_storage().data_ptr() == c.untyped_storage().data_ptr() and m.stride() == c.stride()
                and m.dtype == c.dtype and m.shape[0] == c.shape[0] and m.shape[2:] == c.shape[2:])
        if same:
            sc = m.stride(1)
            d = c.storage_offset() - m.storage_offset()
            if sc > 0 and d % sc == 0 and d // sc >= m.shape[1]:
                n_ch = d // sc + c.shape[1]
                base = torch.as_strided(m, (m.shape[0], n_ch, m.shape[2], m.shape[3]), m.stride(), m.storage_offset())
                return base, 0, d // sc
        return torch.cat([m, c], 1), 0, m.shape[1]      # separate tensors: one device-side copy

    # ------------------------------------------------------------------ RANSAC pieces
    def make_samples(self, counts, cap, H=150, m=5, sampler="cv2", seed=0):
        B = counts.shape[0]
        s = torch.empty((B, H, m), dtype=torch.int32, device=self.device)
        rc = self.lib.zp_make_samples(self.ctx.handle, _ptr(counts), int(cap), B, H, m, _lib.SAMPLER[sampler],
                                      int(seed), _ptr(s), _stream(self.device))
        self.ctx.check(rc, "zp_make_samples")
        return s

    def _Ks(self, Ks, B):
        if isinstance(Ks, torch.Tensor) and Ks.device == self.device and Ks.dtype == torch.float64 and Ks.is_contiguous() \
                and tuple(Ks.shape) == (B, 9):
            return Ks
        K = torch.as_tensor(Ks).to(device=self.device, dtype=torch.float64)
        if K.numel() == 9:
            K = K.reshape(1, 9).expand(B, 9)
        return K.reshape(B, 9).contiguous()

    def solve_minimal(self, corr, counts, Ks, samples):
        B, _, cap = corr.shape
        _, H, m = samples.shape
        K = self._Ks(Ks, B)
        hp = torch.empty((B, H, 12), dtype=torch.float64, device=self.device)
        rc = self.lib.zp_solve_minimal(self.ctx.handle, _ptr(corr), cap, _ptr(counts), _ptr(K), _ptr(samples), B, H, m,
                                       _ptr(hp), _stream(self.device))
        self.ctx.check(rc, "zp_solve_minimal")
        return hp

    def score(self, corr, counts, Ks, hyp_poses, thr=2.0):
        B, _, cap = corr.shape
        H = hyp_poses.shape[1]
        K = self._Ks(Ks, B)
        hp = torch.as_tensor(hyp_poses).to(device=self.device, dtype=torch.float64).contiguous()
        out = torch.empty((B, H), dtype=torch.int32, device=self.device)
        rc = self.lib.zp_score(self.ctx.handle, _ptr(corr), cap, _ptr(counts), _ptr(K), _ptr(hp), B, H, float(thr),
                               _ptr(out), _stream(self.device))
        self.ctx.check(rc, "zp_score")
        return out

    def ransac(self, corr, counts, Ks, *, samples=None, H=150, m=5, thr=2.0, conf=0.99, sampler="cv2", seed=0,
               select="cv2_replay", final="epnp", return_details=False):
        """Returns dict(poses f64 [B,12], n_inliers i32 [B], status i32 [B] [, hyp_poses, hyp_inliers, best_idx,
        inlier_mask])."""
        B, _, cap = corr.shape
        if samples is not None:
            H, m = samples.shape[1], samples.shape[2]
        K = self._Ks(Ks, B)
        out = dict(poses=torch.empty((B, 12), dtype=torch.float64, device=self.device),
                   n_inliers=torch.empty((B,), dtype=torch.int32, device=self.device),
                   status=torch.empty((B,), dtype=torch.int32, device=self.device))
        hp = hi = bi = im = ir = None
        if return_details:           # "hyps": also the hypothesis lists, which forces one wave of all H hypotheses
            bi = out["best_idx"] = torch.empty((B,), dtype=torch.int32, device=self.device)
            ir = out["iters_run"] = torch.empty((B,), dtype=torch.int32, device=self.device)
            im = out["inlier_mask"] = torch.empty((B, cap), dtype=torch.uint8, device=self.device)
            if return_details != "state":
                hp = out["hyp_poses"] = torch.empty((B, H, 12), dtype=torch.float64, device=self.device)
                hi = out["hyp_inliers"] = torch.empty((B, H), dtype=torch.int32, device=self.device)
        rc = self.lib.zp_ransac(self.ctx.handle, _ptr(corr), cap, _ptr(counts), _ptr(K), _ptr(samples), B, int(H),
                                int(m), float(thr), float(conf), _lib.SAMPLER[sampler], int(seed),
                                _lib.SELECT[select], _lib.FINAL[final], _ptr(hp), _ptr(hi), _ptr(bi), _ptr(ir), _ptr(im),
                                _ptr(out["poses"]), _ptr(out["n_inliers"]), _ptr(out["status"]), _stream(self.device))
        self.ctx.check(rc, "zp_ransac")
        return out

    # ------------------------------------------------------------------ the batched entry (what the bench times)
    def decode_and_pose_batch(self, logits, bboxes, Ks, obj_ids=None, *, obj_default=0, mask_ch=0, bit0_ch=1,
                              n_bits=16, ignore_bit=0, ext_mask=None, m=5, iters=150, thr=2.0, conf=0.99,
                              sampler="cv2", seed=0, select="cv2_replay", final="epnp", graph=False, out=None, records=None):
        """Device logits in, device poses out, no host copy, ONE C call (zp_pose_batch_device): poses f64 [B,12]
        (R row-major | t mm), n_inliers i32 [B], status i32 [B] (0 ok, 1 no mask pixel, 2 < 6 correspondences, 3 RANSAC found
        no model).  out = (poses, n_inliers, status) to write into caller-owned tensors; records = f64 [B,14] to also get
        the packed pose | n_inliers | status record of the multi-GPU gather.
        graph=True replays the chain from a CUDA graph: the first call with a given set of tensors captures it, so the
        INPUT AND OUTPUT TENSORS MUST BE THE SAME OBJECTS from call to call (with out=None the engine then returns its own
        persistent output tensors for this batch size, overwritten by the next call)."""
        if isinstance(logits, (tuple, list)):
            logits, mask_ch, bit0_ch = self._join_views(*logits)
        if logits.dim() != 4 or logits.shape[2] != logits.shape[3]:
            raise ValueError("logits must be [B,C,S,S]")
        if logits.dtype not in _DT:
            raise TypeError("logits dtype %s not supported (float32 | bfloat16)" % logits.dtype)
        if logits.device != self.device:
            raise ValueError("logits live on %s, engine on %s" % (logits.device, self.device))
        B, Cc, S, _ = logits.shape
        if bit0_ch + (n_bits - ignore_bit) > Cc or mask_ch >= Cc:
            raise ValueError("channel layout exceeds the %d channels of logits" % Cc)
        bb = self._dev_f64(bboxes, (B, 4))
        K = self._Ks(Ks, B)
        oid = None
        if obj_ids is not None:
            oid = obj_ids if (isinstance(obj_ids, torch.Tensor) and obj_ids.device == self.device and obj_ids.dtype == torch.int32
                              and obj_ids.is_contiguous()) else torch.as_tensor(obj_ids).to(device=self.device, dtype=torch.int32).contiguous()
        em = None
        if ext_mask is not None:
            em = (torch.as_tensor(ext_mask).to(device=self.device) != 0).to(torch.uint8).contiguous().reshape(B, S, S)
        if out is None:
            if graph:
                if B not in self._dev_out:
                    self._dev_out[B] = (torch.empty((B, 12), dtype=torch.float64, device=self.device),
                                        torch.empty((B,), dtype=torch.int32, device=self.device),
                                        torch.empty((B,), dtype=torch.int32, device=self.device))
                out = self._dev_out[B]
            else:
                out = (torch.empty((B, 12), dtype=torch.float64, device=self.device),
                       torch.empty((B,), dtype=torch.int32, device=self.device),
                       torch.empty((B,), dtype=torch.int32, device=self.device))
        poses, ninl, status = out
        strides = (C.c_int64 * 4)(*logits.stride())
        rc = self.lib.zp_pose_batch_device(
            self.ctx.handle, _ptr(logits), _DT[logits.dtype], B, S, strides, int(mask_ch), int(bit0_ch), int(n_bits),
            int(ignore_bit), _ptr(em), _ptr(bb), _ptr(K), _ptr(oid), int(obj_default), int(iters), int(m), float(thr),
            float(conf), _lib.SAMPLER[sampler], int(seed), _lib.SELECT[select], _lib.FINAL[final], _ptr(poses), _ptr(ninl),
            _ptr(status), _ptr(records), 1 if graph else 0, _stream(self.device))
        self.ctx.check(rc, "zp_pose_batch_device")
        if graph:                          # the graph names these buffers: keep them alive as long as the engine
            self._graph_refs[(logits.data_ptr(), bb.data_ptr(), K.data_ptr())] = (logits, bb, K, oid, em, out, records)
        return poses, ninl, status

    def _dev_f64(self, x, shape):
        """float64 device tensor of `shape`; a tensor that already is one is passed through untouched (stable pointer)"""
        if isinstance(x, torch.Tensor) and x.device == self.device and x.dtype == torch.float64 and x.is_contiguous() \
                and tuple(x.shape) == tuple(shape):
            return x
        return torch.as_tensor(x).to(device=self.device, dtype=torch.float64).contiguous().reshape(shape)

    def pose_batch_host(self, logits, bboxes, Ks, obj_ids=None, *, obj_default=0, mask_ch=0, bit0_ch=1, n_bits=16,
                        ignore_bit=0, m=5, iters=150, thr=2.0, conf=0.99, sampler="cv2", seed=0, select="cv2_replay",
                        final="epnp", out=None, asynchronous=False):
        """HOST numpy / pinned-tensor buffers in, HOST results out, through zp_pose_batch_host (H2D + chain + D2H
        inside one C call; synchronous).  logits [B,C,S,S] contiguous float32 (numpy) or a CPU torch tensor.
        asynchronous=True enqueues only (zp_pose_batch_host_async): `out` is valid after sync(); use pinned buffers."""
        lg = logits if isinstance(logits, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(logits))
        if lg.is_cuda or not lg.is_contiguous():
            raise ValueError("pose_batch_host takes contiguous HOST logits")
        B, Cc, S, _ = lg.shape
        if not (0 <= mask_ch < Cc) or bit0_ch < 0 or bit0_ch + (n_bits - ignore_bit) > Cc:
            raise ValueError("channel layout exceeds the %d channels of logits" % Cc)
        bb = np.ascontiguousarray(np.asarray(bboxes, np.float64).reshape(B, 4))
        K = np.asarray(Ks, np.float64)
        K = np.ascontiguousarray(np.broadcast_to(K.reshape(-1, 9), (B, 9)))
        oid = None if obj_ids is None else np.ascontiguousarray(np.asarray(obj_ids, np.int32))
        if out is None:
            out = (np.empty((B, 12)), np.empty(B, np.int32), np.empty(B, np.int32))
        poses, ninl, status = out
        fn = self.lib.zp_pose_batch_host_async if asynchronous else self.lib.zp_pose_batch_host
        if asynchronous:                      # the C side reads these after we return: keep them alive until sync()
            self._inflight.append((lg, bb, K, oid, out))
        rc = fn(
            self.ctx.handle, C.c_void_p(lg.data_ptr()), _DT[lg.dtype], B, Cc, S, int(mask_ch), int(bit0_ch), int(n_bits),
            int(ignore_bit), bb.ctypes.data_as(C.c_void_p), K.ctypes.data_as(C.c_void_p),
            oid.ctypes.data_as(C.c_void_p) if oid is not None else C.c_void_p(), int(obj_default), int(iters), int(m),
            float(thr), float(conf), _lib.SAMPLER[sampler], int(seed), _lib.SELECT[select], _lib.FINAL[final],
            poses.ctypes.data_as(C.c_void_p), ninl.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p))
        self.ctx.check(rc, "zp_pose_batch_host")
        return poses, ninl, status

    def sync(self):
        """waits for an asynchronous pose_batch_host submission"""
        self.ctx.check(self.lib.zp_sync(self.ctx.handle), "zp_sync")
        self._inflight = []

    # ------------------------------------------------------------------ small stand-alone helpers
    def remap_pixels(self, pixels, bbox, S):
        px = torch.as_tensor(np.ascontiguousarray(pixels, dtype=np.int64)).to(self.device)
        out = torch.empty_like(px)
        bb = np.ascontiguousarray(np.asarray(bbox, np.float64).reshape(4))
        rc = self.lib.zp_remap_pixels(self.ctx.handle, _ptr(px), px.shape[0], bb.ctypes.data_as(C.c_void_p), int(S),
                                      _ptr(out), _stream(self.device))
        self.ctx.check(rc, "zp_remap_pixels")
        return out

    def codes_to_ids(self, bits, base=2):
        b = torch.as_tensor(np.ascontiguousarray(bits, dtype=np.float64)).to(self.device)
        N, L = b.shape
        out = torch.empty((N,), dtype=torch.float64, device=self.device)
        rc = self.lib.zp_codes_to_ids(self.ctx.handle, _ptr(b), N, L, int(base), _ptr(out), _stream(self.device))
        self.ctx.check(rc, "zp_codes_to_ids")
        return out

    # ------------------------------------------------------------------ fused network tail
    def upload_head(self, weight, bias=None):
        """Weights of the network's last 1x1 convolution (model/aspp.py:58 conv_1x1_4): weight [n_out, c_in] or the
        Conv2d's [n_out, c_in, 1, 1]; bias [n_out] or None.  Rounded to bfloat16 by the library.  Synchronises."""
        w = torch.as_tensor(weight).detach().to("cpu", torch.float32)
        w = w.reshape(w.shape[0], -1).contiguous().numpy()
        b = None if bias is None else np.ascontiguousarray(torch.as_tensor(bias).detach().to("cpu", torch.float32).numpy())
        rc = self.lib.zp_upload_head(self.ctx.handle, w.ctypes.data_as(C.c_void_p),
                                     b.ctypes.data_as(C.c_void_p) if b is not None else C.c_void_p(),
                                     int(w.shape[0]), int(w.shape[1]))
        self.ctx.check(rc, "zp_upload_head")
        self._head = (int(w.shape[0]), int(w.shape[1]))

    @staticmethod
    def _channels_last_ptr(t, name):
        if t.dim() != 4 or t.dtype not in _DT:
            raise TypeError("%s must be a 4-D bfloat16 | float32 tensor [B,C,S,S] in channels_last memory format" % name)
        if not t.permute(0, 2, 3, 1).is_contiguous():
            raise ValueError("%s must be channels_last (x.contiguous(memory_format=torch.channels_last))" % name)
        return t

    def head_decode(self, x, x_skip, bboxes, obj_ids=None, *, obj_default=0, mask_ch=0, bit0_ch=1, n_bits=16, ignore_bit=0,
                    return_codes=False, cap=None):
        """`conv_1x1_4(torch.cat([x, x_skip], 1))` (model/aspp.py:112) + decode() in one pass on the tensor cores; the
        logits are never written.  x [B,c1,S,S], x_skip [B,c2,S,S] | None: bfloat16 (or float32: TF32 tensor-core products),
        channels_last.  Returns like decode()."""
        x = self._channels_last_ptr(x, "x")
        B, c1, S, S2 = x.shape
        c2 = 0
        if x_skip is not None:
            x_skip = self._channels_last_ptr(x_skip, "x_skip")
            if x_skip.shape[0] != B or x_skip.shape[2:] != x.shape[2:]:
                raise ValueError("x and x_skip must agree in batch and spatial size")
            c2 = x_skip.shape[1]
            if x_skip.dtype != x.dtype:
                raise TypeError("x and x_skip must have the same dtype")
        if S != S2 or x.device != self.device:
            raise ValueError("x must be [B,C,S,S] on %s" % self.device)
        cap = int(cap or ((S * S + 3) // 4) * 4)
        bb = torch.as_tensor(bboxes).to(device=self.device, dtype=torch.float64).contiguous().reshape(B, 4)
        oid = None
        if obj_ids is not None:
            oid = torch.as_tensor(obj_ids).to(device=self.device, dtype=torch.int32).contiguous()
        corr = torch.empty((B, 5, cap), dtype=torch.float32, device=self.device)
        counts = torch.empty((B,), dtype=torch.int32, device=self.device)
        codes = torch.empty((B, S, S), dtype=torch.uint16, device=self.device) if return_codes else None
        rc = self.lib.zp_head_decode(self.ctx.handle, _ptr(x), int(c1), _ptr(x_skip), int(c2), _DT[x.dtype], B, S, int(mask_ch),
                                     int(bit0_ch), int(n_bits), int(ignore_bit), _ptr(bb), _ptr(oid), int(obj_default),
                                     _ptr(codes), _ptr(corr), cap, _ptr(counts), _stream(self.device))
        self.ctx.check(rc, "zp_head_decode")
        return (corr, counts, codes) if return_codes else (corr, counts)

    def head_pose_batch(self, x, x_skip, bboxes, Ks, obj_ids=None, *, obj_default=0, mask_ch=0, bit0_ch=1, n_bits=16,
                        ignore_bit=0, m=5, iters=150, thr=2.0, conf=0.99, sampler="cv2", seed=0, select="cv2_replay",
                        final="epnp"):
        """decode_and_pose_batch() fed by the network's last activations instead of its logits (head_decode + ransac)."""
        corr, counts = self.head_decode(x, x_skip, bboxes, obj_ids, obj_default=obj_default, mask_ch=mask_ch,
                                        bit0_ch=bit0_ch, n_bits=n_bits, ignore_bit=ignore_bit)
        r = self.ransac(corr, counts, Ks, H=iters, m=m, thr=thr, conf=conf, sampler=sampler, seed=seed, select=select,
                        final=final)
        return r["poses"], r["n_inliers"], r["status"]

    # ------------------------------------------------------------------ either side of the path
    def final_bboxes(self, det_boxes, padding_ratio=1.5, resize_method="crop_square_resize", max_x=640, max_y=480):
        """padding_Bbox + get_final_Bbox (bop_dataset_pytorch.py:123-139, 162-194) for [B,4] detection boxes (x,y,w,h)
        on the device.  padding_ratio <= 0 skips padding_Bbox.  Returns float64 [B,4] with integral values, the
        `bboxes` argument of decode()."""
        bb = torch.as_tensor(det_boxes).to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 4)
        out = torch.empty_like(bb)
        rc = self.lib.zp_final_bbox(self.ctx.handle, _ptr(bb), bb.shape[0], float(padding_ratio),
                                    _lib.RESIZE[resize_method], float(max_x), float(max_y), _ptr(out), _stream(self.device))
        self.ctx.check(rc, "zp_final_bbox")
        return out

    def crop_inputs(self, images, padded_boxes, img_ids=None, *, crop_size=256, resize_method="crop_square_resize",
                    dtype=torch.float32, channels_last=False, return_u8=False):
        """get_roi(INTER_LINEAR) + ToTensor + Normalize (bop_dataset_pytorch.py:110-121, 334-347) for B boxes at once:
        images uint8 [n_img,H,W,3] | [H,W,3] on the device, padded_boxes [B,4] (output of padding_Bbox), img_ids [B] | None.
        Returns the network input [B,3,cs,cs] (float32 | bfloat16; channels_last memory format if asked) (, uint8 crops)."""
        im = torch.as_tensor(images)
        if im.dim() == 3:
            im = im.unsqueeze(0)
        if im.dtype != torch.uint8 or im.dim() != 4 or im.shape[3] != 3:
            raise TypeError("images must be uint8 [n_img,H,W,3]")
        im = im.to(self.device).contiguous()
        bb = torch.as_tensor(padded_boxes).to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 4)
        B = bb.shape[0]
        ids = None if img_ids is None else torch.as_tensor(img_ids).to(device=self.device, dtype=torch.int32).contiguous()
        cs = int(crop_size)
        out = torch.empty((B, 3, cs, cs), dtype=dtype, device=self.device,
                          memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        u8 = torch.empty((B, cs, cs, 3), dtype=torch.uint8, device=self.device) if return_u8 else None
        rc = self.lib.zp_crop_input(self.ctx.handle, _ptr(im), im.shape[0], im.shape[1], im.shape[2], _ptr(ids), _ptr(bb), B,
                                    cs, _lib.RESIZE[resize_method], C.c_void_p(), C.c_void_p(), _DT[dtype],
                                    1 if channels_last else 0, _ptr(out), _ptr(u8), _stream(self.device))
        self.ctx.check(rc, "zp_crop_input")
        return (out, u8) if return_u8 else out

    def upload_model(self, obj_id, vertices):
        """Model vertices [V,3] (mm) of object slot obj_id for pose_errors().  Synchronises."""
        v = np.ascontiguousarray(np.asarray(vertices, np.float64).reshape(-1, 3))
        rc = self.lib.zp_upload_model(self.ctx.handle, int(obj_id), v.ctypes.data_as(C.c_void_p), int(v.shape[0]))
        self.ctx.check(rc, "zp_upload_model")

    def pose_errors(self, poses_est, poses_gt, obj_ids=None, *, obj_default=0, add=True, adi=True):
        """ADD / ADI (lib/pysixd/pose_error.py:297-336) of B pose pairs, poses float64 [B,12] (R row-major | t mm).
        Returns (add [B] | None, adi [B] | None), float64 on the device."""
        pe = torch.as_tensor(poses_est).to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 12)
        pg = torch.as_tensor(poses_gt).to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 12)
        if pe.shape != pg.shape:
            raise ValueError("poses_est and poses_gt must have the same shape")
        B = pe.shape[0]
        oid = None
        if obj_ids is not None:
            oid = torch.as_tensor(obj_ids).to(device=self.device, dtype=torch.int32).contiguous()
        o_add = torch.empty((B,), dtype=torch.float64, device=self.device) if add else None
        o_adi = torch.empty((B,), dtype=torch.float64, device=self.device) if adi else None
        rc = self.lib.zp_pose_errors(self.ctx.handle, _ptr(pe), _ptr(pg), _ptr(oid), int(obj_default), B,
                                     _ptr(o_add), _ptr(o_adi), _stream(self.device))
        self.ctx.check(rc, "zp_pose_errors")
        return o_add, o_adi

    def launch_count(self):
        return int(self.lib.zp_launch_count(self.ctx.handle))

    def set_kernel_timing(self, on=True):
        """bracket every main-kernel launch with CUDA events on its stream (serialises; measurement passes only)"""
        self.ctx.check(self.lib.zp_set_kernel_timing(self.ctx.handle, 1 if on else 0), "zp_set_kernel_timing")

    def kernel_time(self, name):
        """(average ms per launch, launches) of kernel `name` since set_kernel_timing(True)"""
        ms, n = C.c_double(), C.c_int64()
        self.ctx.check(self.lib.zp_kernel_time(self.ctx.handle, name.encode(), C.byref(ms), C.byref(n)), "zp_kernel_time")
        return (ms.value / n.value if n.value else 0.0), int(n.value)

    def fp64_peak_tflops(self, iters=10000):
        """measured FP64 FMA throughput of this GPU (DFMA chains, 2 flop per instruction)"""
        v = C.c_double()
        self.ctx.check(self.lib.zp_fp64_peak_probe(self.ctx.handle, int(iters), C.byref(v)), "zp_fp64_peak_probe")
        return v.value

    def fp32_peak_tflops(self, iters=20000, packed=False):
        """measured FP32 FMA throughput of this GPU: scalar FFMA chains, or packed FFMA2 chains (packed=True)"""
        v = C.c_double()
        fn = self.lib.zp_fp32x2_peak_probe if packed else self.lib.zp_fp32_peak_probe
        self.ctx.check(fn(self.ctx.handle, int(iters), C.byref(v)), "zp_fp32_peak_probe")
        return v.value


class Pipeline:
    """Several engines ("lanes": one zp_ctx + one CUDA stream each) used round-robin, so that consecutive batches
    overlap on the GPU: the kernels of this path are latency-bound at BASELINE's 64-crop batches (a step keeps well
    under half of the SMs' issue slots busy), and the host->device copy of one batch runs under the kernels of the
    previous one.  Every lane holds its own copy of the dictionaries and its own workspace."""

    def __init__(self, device=None, lanes=6):
        self.engines = [Engine(device) for _ in range(int(lanes))]
        self.device = self.engines[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.engines]
        self._next = 0
        self._busy = [False] * len(self.engines)

    @property
    def next_lane(self):
        return self._next

    def upload_dict(self, *args, **kw):
        for e in self.engines:
            e.upload_dict(*args, **kw)

    def submit(self, logits, bboxes, Ks, obj_ids=None, post=None, **kw):
        """device-resident batch on the next lane; returns (poses, n_inliers, status, done_event).  The inputs must have
        been produced on (or be visible to) the current stream: the lane's stream waits for it first.  `post`, if given,
        maps the result tuple inside the lane's stream context (e.g. the multi-GPU pose gather)."""
        i = self._next
        self._next = (i + 1) % len(self.engines)
        s = self.streams[i]
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            out = self.engines[i].decode_and_pose_batch(logits, bboxes, Ks, obj_ids, **kw)
            if post is not None:
                out = tuple(post(out))
            ev = torch.cuda.Event()
            ev.record(s)
        for t in (logits, bboxes, Ks, obj_ids):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(s)
        return out + (ev,)

    def join(self):
        """makes the current stream wait for every lane"""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def submit_host(self, logits, bboxes, Ks, obj_ids=None, **kw):
        """host buffers on the next lane (asynchronous); returns the lane index to pass to wait_host().  A lane that
        still has a submission in flight is waited for first."""
        i = self._next
        self._next = (i + 1) % len(self.engines)
        if self._busy[i]:
            self.wait_host(i)
        self.engines[i].pose_batch_host(logits, bboxes, Ks, obj_ids, asynchronous=True, **kw)
        self._busy[i] = True
        return i

    def wait_host(self, lane=None):
        for i in (range(len(self.engines)) if lane is None else [lane]):
            if self._busy[i]:
                self.engines[i].sync()
                self._busy[i] = False

    def launch_count(self):
        return sum(e.launch_count() for e in self.engines)


_default = {}


def default_engine(device=None):
    """Process-wide engine per device, created on first use (what the drop-in functions call)."""
    idx = torch.cuda.current_device() if device is None else device
    if idx not in _default:
        _default[idx] = Engine(idx)
    return _default[idx]
