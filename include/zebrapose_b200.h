/*
 * zebrapose_b200 -- C ABI of the B200-native post-network pose path of ZebraPose.
 *
 * The reference (lyltc1/ZebraPose) has no FFI for this path: the boundary is plain Python
 * (zebrapose/binary_code_helper/CNN_output_to_pose.py, class_id_encoder_decoder.py,
 * generate_new_dict.py, zebrapose/common_ops.py; call sites zebrapose/test.py:250-273,
 * zebrapose/test_vivo.py:160-172).  This header is what a ctypes binding of that path binds
 * (see INTEGRATION.md); zebrapose_b200/_lib.py is exactly that binding.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller (e.g. a torch tensor's data_ptr())
 *     unless its comment says "host".  The library allocates nothing the caller must free except
 *     the context (internal workspace lives and dies with it).
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *     no hidden synchronisation except where stated ("synchronises").
 *   - Return value: 0 = ok, < 0 = error; zp_last_error(ctx) describes the last failure.
 *   - One zp_ctx per (device, host thread).  Not thread-safe across threads sharing a ctx.
 *   - Correspondence lists are SoA per crop: float corr[B][5][cap], planes u, v, X, Y, Z
 *     (u,v = original-image pixel, float32 of an integer; X,Y,Z = model point in mm), entries
 *     [0, counts[b]) valid, in row-major pixel order of the crop (CNN_output_to_pose.py:111,54).
 *   - Poses are double[12] per crop/hypothesis: R row-major (9) then t (3, millimetres).
 */
#ifndef ZEBRAPOSE_B200_H
#define ZEBRAPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zp_ctx zp_ctx;

enum { ZP_DTYPE_F32 = 0, ZP_DTYPE_BF16 = 1 };
/* non-existing (NaN) dictionary rows: ZERO = reference behaviour (3D point (0,0,0), pixel kept,
 * CNN_output_to_pose.py:58-62); HAMMING = north_star extension (nearest existing code). */
enum { ZP_NONEXIST_ZERO = 0, ZP_NONEXIST_HAMMING = 1 };
enum { ZP_SAMPLER_CV2 = 0, ZP_SAMPLER_PHILOX = 1 };
enum { ZP_SELECT_CV2_REPLAY = 0, ZP_SELECT_ARGMAX = 1 };
enum { ZP_FINAL_EPNP = 0, ZP_FINAL_EPNP_GN = 1 };
/* minimal solver of the RANSAC hypotheses: CV2 = operation-by-operation replay of OpenCV's EPnP (hypotheses bit-identical
 * to cv2.solvePnP(SOLVEPNP_EPNP), so the RANSAC winner is cv2's); FAST = float64 EPnP with a bisection / inverse-iteration
 * null space (accurate, ~1e-5 deg from cv2 for m >= 6, but for 4/5-point samples it returns another null-space basis than
 * cv2 and therefore other hypotheses). */
enum { ZP_SOLVER_CV2 = 0, ZP_SOLVER_FAST = 1 };
/* per-crop status written by zp_ransac */
enum { ZP_OK = 0, ZP_NO_MASK_PIXELS = 1, ZP_TOO_FEW_POINTS = 2, ZP_RANSAC_NO_MODEL = 3 };

/* resize_method of get_final_Bbox (bop_dataset_pytorch.py:169,184; config key `resize_method`) */
/* ZP_CROP_KEEP: no final step (padding_Bbox only; get_final_Bbox returns the box unchanged for any other string). */
enum { ZP_CROP_RESIZE = 0, ZP_CROP_SQUARE_RESIZE = 1, ZP_CROP_RESIZE_BY_WARP_AFFINE = 2, ZP_CROP_KEEP = 3 };

#define ZP_MAX_OBJECTS 256
#define ZP_MAX_HYPOTHESES 1024

int zp_version(void);
int zp_create(zp_ctx** out, int device);
void zp_destroy(zp_ctx* ctx);
const char* zp_last_error(zp_ctx* ctx);

/* Replaces load_dict_class_id_3D_points' table + generate_new_corres_dict (generate_new_dict.py:4-33):
 * uploads the correspondence dictionary of object slot `obj_id` (0 <= obj_id < ZP_MAX_OBJECTS).
 * pts_xyz: HOST double [2^n_bits][3], NaN rows = non-existing codes.  For ignore_bit = k the parent table
 * (2^(n_bits-k) rows) is built on the host in the reference's order (float64, children summed in ascending
 * id order, then / 2^k; NaN propagates) and cast to float32 like CNN_output_to_pose.py:129.  Synchronises. */
int zp_upload_tables(zp_ctx* ctx, int obj_id, const double* pts_xyz, int n_bits, int ignore_bit,
                     int nonexist_mode);
/* Debug/parity read-back of what zp_upload_tables built (HOST outputs, both nullable):
 * pts_out float[2^(n_bits-k)][4] (x,y,z,exists), remap_out uint16[2^(n_bits-k)].  Synchronises. */
int zp_download_tables(zp_ctx* ctx, int obj_id, float* pts_out, uint16_t* remap_out);

/* Replaces common_ops.from_output_to_class_mask / from_output_to_class_binary_code (common_ops.py:5-19),
 * class_code_images_to_class_id_image (class_id_encoder_decoder.py:17-28), build_non_unique_2D_3D_correspondence
 * (CNN_output_to_pose.py:53-64) and mapping_pixel_position_to_original_position (:34-50), for B crops at once.
 *   logits   [B, C, S, S] network output, any strides (elements): strides[4] = {batch, channel, row, pixel}
 *   mask_ch  channel of the mask logit (0), bit0_ch first code-bit channel (1; v2 nets: 2), MSB first
 *   n_bits   code length (16); ignore_bit k: only the first n_bits-k bit planes are read (test.py:267)
 *   ext_mask nullable uint8 [B,S,S]; non-zero = masked; overrides mask_ch (mask-rcnn variants)
 *   bbox     double [B,4] = x, y, w, h of get_final_Bbox (bop_dataset_pytorch.py:162-194)
 *   obj_ids  nullable int32 [B] table slot per crop; NULL -> obj_default for all
 *   codes    nullable uint16 [B,S,S] out: class id of every pixel
 *   corr     float [B,5,cap] out; counts int32 [B] out (number of masked pixels, may exceed cap -> clipped lists)
 */
int zp_decode(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4],
              int mask_ch, int bit0_ch, int n_bits, int ignore_bit, const uint8_t* ext_mask,
              const double* bbox, const int32_t* obj_ids, int obj_default,
              uint16_t* codes, float* corr, int cap, int32_t* counts, void* stream);

/* zp_decode for the CE heads of the ablation configs (common_ops.py:21-30 + class_code_images_to_class_id_image with
 * class_base = `base`, class_id_encoder_decoder.py:17-28): the code logits are n_digits groups of `base` consecutive
 * channels starting at digit0_ch; a digit is the first maximum of the float32 softmax of its group, the class id the
 * base-`base` number with digit 0 most significant.  base^n_digits <= 65536; the slot's table must hold at least that
 * many rows with ignore_bit 0.  Everything else as zp_decode.  (A digit can differ from the reference's only where
 * float32 softmax rounding creates or breaks an exact tie differently in torch's exp than in CUDA's.) */
int zp_decode_ce(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4],
                 int mask_ch, int digit0_ch, int base, int n_digits, const uint8_t* ext_mask,
                 const double* bbox, const int32_t* obj_ids, int obj_default,
                 uint16_t* codes, float* corr, int cap, int32_t* counts, void* stream);

/* Stand-alone device forms of two small reference helpers (the batched path has them fused into zp_decode):
 * mapping_pixel_position_to_original_position (CNN_output_to_pose.py:34-50): px int64 [N,2] (x,y) -> out int64 [N,2];
 * h_bbox is a HOST double[4].  class_code_images_to_class_id_image (class_id_encoder_decoder.py:17-28):
 * bits double [N,L] (values 0..base-1, channel 0 most significant) -> ids double [N] = sum bits[i]*base^(L-1-i). */
int zp_remap_pixels(zp_ctx* ctx, const int64_t* px, int64_t N, const double* h_bbox, int S, int64_t* out, void* stream);
int zp_codes_to_ids(zp_ctx* ctx, const double* bits, int64_t N, int L, int base, double* ids, void* stream);

/* Minimal-sample index lists, exportable for parity runs.  mode CV2: replays cv::RNG(0xFFFFFFFFFFFFFFFF)
 * exactly as cv2.solvePnPRansac draws them (depends on counts[b] only); PHILOX: counter-based, seeded.
 * samples int32 [B,H,m] out; crops with counts[b] < m get -1. */
int zp_make_samples(zp_ctx* ctx, const int32_t* counts, int cap, int B, int H, int m, int mode,
                    uint64_t seed, int32_t* samples, void* stream);

/* Minimal solver used by zp_solve_minimal / zp_ransac / zp_pose_batch_host (ZP_SOLVER_*; default CV2). */
int zp_set_solver(zp_ctx* ctx, int solver);
/* Wave plan of zp_ransac: hypotheses are solved and scored in waves of sizes[0], sizes[1], ... (HOST int32[n], n <= 16;
 * the last size repeats until H is covered) and after every wave cv2's adaptive-stop rule is replayed, so crops that
 * have reached their stopping iteration skip the remaining waves.  n = 0: automatic.  Results do not depend on the plan
 * (cv2 never consults a hypothesis at or past its stopping iteration). */
int zp_set_waves(zp_ctx* ctx, int n, const int32_t* sizes);

/* Near-ties in the replay of cv2's update rule (`good > maxGood`, CNN_output_to_pose.py:155-157 -> cv2.solvePnPRansac): on = 1
 * (default) re-counts, with cv2's own double -> float32 arithmetic for the points the FP32 scoring predicate puts within
 * 1e-3 px of the threshold, every hypothesis that comes within 3 inliers of the running maximum (and, once, the record
 * holder), so those decisions are taken on cv2's counts; on = 0 lets the FP32 counts decide everything (they can differ by
 * one for such a point, the slack north_star allows); on = 2 (test aid) re-counts every near-tie at once instead of parking
 * the early low-count ones -- the rule that parking must reproduce. */
int zp_set_exact_ties(zp_ctx* ctx, int on);

/* Shape of the final solve on the winner's inliers: 2 = split into three kernels (point moments over 4 CTAs per crop ->
 * a warp per crop for the solver chain -> candidate errors + pick; EPnP's sums as contractions of raw moments), 4 = a
 * thread-block cluster of four CTAs per crop (partial sums combined through distributed shared memory), 1 = one CTA per
 * crop walking the same four point partitions in turn, 0 = automatic (split; with final = "epnp+gn" the cluster while 4 B
 * CTAs fit one wave, else one CTA).  Forms 1 and 4 produce identical bits; the split form agrees with them to rounding
 * (1e-9 deg / 1e-9 mm on BASELINE's crops), and none of the forms depends on the batch a crop came in. */
int zp_set_final_form(zp_ctx* ctx, int form);

/* EPnP on each m-point minimal set (float64).  K double [B,9] row-major.  hyp_poses double [B,H,12] out;
 * hypotheses of crops with too few points or degenerate samples are written as NaN. */
int zp_solve_minimal(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                     const int32_t* samples, int B, int H, int m, double* hyp_poses, void* stream);

/* Reprojection scoring of H hypotheses against all correspondences of each crop (the FP32 kernel):
 * inlier <=> (x - u z)^2 + (y - v z)^2 <= thr^2 z^2 with [x y z] = K [R|t] [X Y Z 1], float32.
 * hyp_inliers int32 [B,H] out.  Replaces PnPRansacCallback::computeError of cv2.solvePnPRansac
 * (called at CNN_output_to_pose.py:155-157). */
int zp_score(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
             const double* hyp_poses, int B, int H, float thr_px, int32_t* hyp_inliers, void* stream);

/* Whole RANSAC-PnP (replaces cv2.solvePnPRansac(..., reprojectionError=thr_px, iterationsCount=H,
 * flags=SOLVEPNP_EPNP) + cv2.Rodrigues, CNN_output_to_pose.py:155-158):
 * samples (nullable: generated internally with `sampler`/`seed`) -> minimal EPnP -> scoring -> winner
 * (CV2_REPLAY: cv2's strictly-greater update with RANSACUpdateNumIters(confidence) replayed over the H counts;
 * ARGMAX: most inliers, lowest index on ties) -> EPnP on the winner's inliers (+ optional Gauss-Newton polish).
 * Outputs: poses double [B,12], n_inliers int32 [B], status int32 [B]; nullable hyp_poses double [B,H,12],
 * hyp_inliers int32 [B,H], best_idx int32 [B], iters_run int32 [B] (iterations cv2's loop runs before its adaptive
 * stop), inlier_mask uint8 [B,cap].  Asking for hyp_poses / hyp_inliers makes one wave of all H hypotheses; otherwise
 * hypotheses at or past a crop's stopping iteration may not be computed (zp_set_waves). */
int zp_ransac(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
              const int32_t* samples, int B, int H, int m, float thr_px, double confidence,
              int sampler, uint64_t seed, int select_mode, int final_mode,
              double* hyp_poses, int32_t* hyp_inliers, int32_t* best_idx, int32_t* iters_run, uint8_t* inlier_mask,
              double* poses, int32_t* n_inliers, int32_t* status, void* stream);

/* The whole path in ONE call on device-resident buffers: zp_decode + zp_ransac with the correspondence lists kept in the
 * context (arguments as for those two; bbox double [B,4], K double [B,9]).  records (nullable): double [B,14] = pose |
 * n_inliers | status, the fixed-size record the multi-GPU gather exchanges, written by the final-solve kernel itself.
 * use_graph != 0: the first call with a given argument set (pointers, shapes, options, stream) runs eagerly and captures
 * the 10+ dependent launches of the chain into a CUDA graph; later calls with the same arguments are one cudaGraphLaunch.
 * The buffers a captured call names must stay allocated while the context lives (or until the arguments change); at most
 * 32 argument sets are cached per context; calls on the legacy default stream (stream = NULL) cannot be captured and stay
 * eager.  Results are identical with and without the graph. */
int zp_pose_batch_device(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4],
                         int mask_ch, int bit0_ch, int n_bits, int ignore_bit, const uint8_t* ext_mask,
                         const double* bbox, const double* K, const int32_t* obj_ids, int obj_default,
                         int H, int m, float thr_px, double confidence, int sampler, uint64_t seed,
                         int select_mode, int final_mode,
                         double* poses, int32_t* n_inliers, int32_t* status, double* records, int use_graph, void* stream);

/* The reference-facing one-call form with HOST buffers (what a per-batch drop-in of test.py:250-273 calls):
 * copies logits/bbox/K/obj_ids host->device, runs decode + RANSAC, copies poses/n_inliers/status back.
 * h_logits: HOST [B,C,S,S] contiguous (dtype as above); h_bbox HOST double [B,4]; h_K HOST double [B,9];
 * h_obj_ids nullable HOST int32 [B].  Outputs HOST: poses double [B,12], n_inliers int32 [B], status int32 [B].
 * Pinned host memory makes the copies asynchronous; the call synchronises before returning. */
int zp_pose_batch_host(zp_ctx* ctx, const void* h_logits, int dtype, int B, int C, int S,
                       int mask_ch, int bit0_ch, int n_bits, int ignore_bit,
                       const double* h_bbox, const double* h_K, const int32_t* h_obj_ids, int obj_default,
                       int H, int m, float thr_px, double confidence, int sampler, uint64_t seed,
                       int select_mode, int final_mode,
                       double* h_poses, int32_t* h_n_inliers, int32_t* h_status);

/* Test / profiling aid: pin zp_decode to one kernel family.  0 = automatic (fused register-staged streaming kernel: a
 * CTA walks several consecutive runs of 2048 fp32 / 4096 bf16 pixels of a crop and prefetches the next run under the
 * emission of the current one; crops of more than 16 runs take the two-kernel path, layouts neither can take the
 * generic kernel), 1 = single-run fused kernel with the compaction bases exchanged through a thread-block cluster
 * (DSMEM), 2 = generic strided kernel, 3 = fused TMA-ring streaming kernel, 4 = two kernels (plane stream -> codes +
 * mask bits, then rank/gather/emit), 6 = single-run fused kernel with independent CTAs; 100 + r = path 0 with r runs
 * per CTA (100 = automatic again).  All produce identical output; DESIGN.md section 4 has the measured comparison. */
int zp_set_decode_path(zp_ctx* ctx, int path);

/* Profiling aid: when set (device pointer to 4 uint64 per decode CTA and at least 16 uint64, or NULL to switch off), the fused
 * decode kernel stores %globaltimer stamps per CTA: [0] start, [1] planes read + ranks known, [3] end
 * (tools/dbg_decode_ctas.py); the exact solver's null-space kernel stores clock64 stamps of its first warp in [0..5]
 * (tools/dbg_null_phases.py) and the split final solve's solver kernel those of crop 0 in [8..13]
 * (tools/dbg_finsolve_phases.py). */
int zp_debug_buffer(zp_ctx* ctx, void* dev_u64);

/* Tuning aid for zp_score / zp_ransac: `groups` = warp-groups (128 threads each) per scoring CTA that split the
 * hypotheses of a work item (0 = automatic = 1, else 1, 2 or 4); `hyp_chunk` = hypotheses per work item (0 = automatic:
 * H/2 when the batch has about one correspondence tile per CTA slot, all of them otherwise; -1 = never cut; else equal
 * chunks of that size).  Results do not depend on either. */
int zp_set_score_groups(zp_ctx* ctx, int groups, int hyp_chunk);

/* Asynchronous form of zp_pose_batch_host: enqueues the copies and the kernels on the ctx's own stream and returns; the
 * outputs are valid after zp_sync(ctx).  The host buffers (inputs AND outputs) must stay alive and unmodified until
 * then and should be pinned (pageable memory makes the copies synchronous).  One submission may be in flight per ctx;
 * several ctxs ("lanes") overlap the host->device copy of one batch with the kernels of another. */
int zp_pose_batch_host_async(zp_ctx* ctx, const void* h_logits, int dtype, int B, int C, int S,
                             int mask_ch, int bit0_ch, int n_bits, int ignore_bit,
                             const double* h_bbox, const double* h_K, const int32_t* h_obj_ids, int obj_default,
                             int H, int m, float thr_px, double confidence, int sampler, uint64_t seed,
                             int select_mode, int final_mode,
                             double* h_poses, int32_t* h_n_inliers, int32_t* h_status);
/* Waits for everything enqueued on the ctx's own stream (zp_pose_batch_host_async). */
int zp_sync(zp_ctx* ctx);

/* Measurement aid: while on, every launch of the path's main kernels is bracketed by two CUDA events recorded on the
 * launching stream directly around it (the call then waits for that kernel), and the elapsed times accumulate per kernel
 * name ("zp_decode_stream_kernel", "zp_samples_kernel", "zp_cvs_prep_kernel" / "_null_" / "_cand_" / "_pick_" (exact solver) or
 * "zp_minimal_kernel" (fast solver), "zp_score_kernel", "zp_rs_replay_kernel", "zp_fin_moments_kernel" / "_solve_" / "_errors_"
 * (split final solve) or "zp_final_kernel" (one-kernel forms),
 * "zp_head_codes_kernel", "zp_decode_emit_kernel", "zp_adi_kernel", "zp_crop_kernel").  Switching it on clears the sums.
 * bench.py's roofline figures are algorithmic bytes (flops) per launch / these durations. */
int zp_set_kernel_timing(zp_ctx* ctx, int on);
int zp_kernel_time(zp_ctx* ctx, const char* kernel_name, double* ms_sum, int64_t* launches);

/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches claim). */
int64_t zp_launch_count(zp_ctx* ctx);

/* Profiling aid: SM-clock timestamps of the phases of CTA 0 of the last RANSAC launches (24 slots: 0-9 final solve,
 * 10-15 minimal solver, 16-19 eigen-solver; synchronises). */
int zp_debug_clocks(zp_ctx* ctx, int64_t* out24);

/* FP32 FMA-chain microbenchmark (the roofline denominator for zp_score has no entry in MEASURED_PEAKS.json):
 * runs `iters` dependent-chain FMAs x 8 chains per thread on the whole chip, returns achieved TFLOP/s in *out.
 * Synchronises. */
int zp_fp32_peak_probe(zp_ctx* ctx, int iters, double* out_tflops);
/* FP64 counterpart (DFMA chains, 2 flop per instruction): the roofline denominator of the EPnP solver kernels.  The exact
 * solver uses unfused multiplies and adds (one flop per instruction), so its ceiling is half this figure. */
int zp_fp64_peak_probe(zp_ctx* ctx, int iters, double* out_tflops);
/* Same with packed FFMA2 (fma.rn.f32x2) chains: the form zp_score_kernel uses. */
int zp_fp32x2_peak_probe(zp_ctx* ctx, int iters, double* out_tflops);

/* ---- the steps either side of the pose path ---- */

/* Replaces padding_Bbox + get_final_Bbox (bop_dataset_pytorch.py:123-139, 162-194; called per detection at
 * test_vivo.py:147-150 and in the dataset's __getitem__) for B detection boxes at once, on the device, so the crop box
 * zp_decode consumes never visits the host.  det_boxes double [B,4] = x, y, w, h; padding_ratio <= 0 skips padding_Bbox;
 * max_x / max_y are the image bounds of the "crop_resize" clamp.  out_boxes double [B,4] (integral values): float64
 * arithmetic with truncation toward zero exactly as the reference's int(). */
int zp_final_bbox(zp_ctx* ctx, const double* det_boxes, int B, double padding_ratio, int resize_method,
                  double max_x, double max_y, double* out_boxes, void* stream);

/* Replaces get_roi(x, Bbox, crop_size_img, cv2.INTER_LINEAR, resize_method) + transforms.ToTensor() + Normalize
 * (bop_dataset_pytorch.py:36-89, 110-121, 303, 334-347; test_vivo.py:152-159): the network's input crops, straight from
 * device-resident images.  images uint8 [n_img,H,W,3] (RGB, contiguous); img_ids nullable int32 [B] (NULL: image 0);
 * boxes double [B,4] = the PADDED box (output of padding_Bbox; get_roi squares / clips it itself), integral values.
 * resize_method ZP_CROP_RESIZE | ZP_CROP_SQUARE_RESIZE.  mean3 / std3 HOST float[3] or NULL (ImageNet constants of the
 * reference).  out: float32 or bfloat16 (out_dtype), [B,3,cs,cs] NCHW or channels-last memory; out_u8 nullable uint8
 * [B,cs,cs,3] = the resized crop before ToTensor.  The uint8 crop and the float32 tensor are bit-identical to the
 * reference's (cv2's fixed-point bilinear resize is restated exactly); bfloat16 is that tensor rounded to nearest even. */
int zp_crop_input(zp_ctx* ctx, const uint8_t* images, int n_img, int H, int W, const int32_t* img_ids,
                  const double* boxes, int B, int crop_size, int resize_method, const float* mean3, const float* std3,
                  int out_dtype, int channels_last, void* out, uint8_t* out_u8, void* stream);

/* Model vertices of object slot `obj_id` for the pose-error metrics: HOST double [V][3] in mm (the `vertices` argument
 * of metric.py:8-18).  Synchronises. */
int zp_upload_model(zp_ctx* ctx, int obj_id, const double* pts_xyz, int V);

/* Replaces Calculate_ADD_Error_BOP / Calculate_ADI_Error_BOP (metric.py:8-18 -> pose_error.add / adi,
 * lib/pysixd/pose_error.py:297-336; evaluated per crop at test.py:465-483) for B pose pairs at once.
 * poses_est / poses_gt double [B,12] (R row-major | t in mm); obj_ids nullable int32 [B] model slot per pair.
 * add_out / adi_out double [B], either may be NULL.  ADD is float64; ADI is the exact nearest neighbour (brute force
 * instead of the reference's cKDTree) on float32 squared distances of points recentred by -t_est. */
int zp_pose_errors(zp_ctx* ctx, const double* poses_est, const double* poses_gt, const int32_t* obj_ids,
                   int obj_default, int B, double* add_out, double* adi_out, void* stream);

/* ---- fused network tail ---- */

/* Weights of the network's last layer, conv_1x1_4 = nn.Conv2d(256 + 64, num_classes, kernel_size=1) (model/aspp.py:58):
 * weight HOST float [n_out][c_in] (the Conv2d weight with its two trailing 1x1 dims dropped), bias HOST float [n_out] or
 * NULL.  n_out <= 32, c_in a multiple of 64 up to 512.  Rounded to bfloat16 (round-to-nearest-even).  Synchronises. */
int zp_upload_head(zp_ctx* ctx, const float* weight, const float* bias, int n_out, int c_in);

/* Replaces `conv_1x1_4(torch.cat([x, x_128], 1))` (model/aspp.py:112) + the whole of zp_decode: the 1x1 convolution runs
 * on the tensor cores (tcgen05, accumulators in TMEM, activations streamed by TMA), its epilogue thresholds and packs
 * the bits, and the correspondence lists are emitted from 2 B/pixel codes -- the logits are never written.
 *   x      [B,S,S,c1] (channels-last: the memory of a torch channels_last [B,c1,S,S] tensor), 16-byte aligned;
 *          dtype ZP_DTYPE_BF16 (kind::f16 MMA; c1, c2 multiples of 64) or ZP_DTYPE_F32 (kind::tf32 MMA: the products use
 *          the operands' top 10 mantissa bits, accumulation is fp32; c1, c2 multiples of 32)
 *   x_skip [B,S,S,c2] or NULL with c2 = 0 (the skip connection the reference concatenates); c1 + c2 = c_in
 *   mask_ch / bit0_ch / n_bits / ignore_bit: output-channel layout as in zp_decode (0 / 1 / 16 / k)
 *   other arguments and outputs exactly as zp_decode.  S*S must be a multiple of 128.
 * A pixel's bit is (sum_c w[o][c] * x[c] accumulated in fp32 + bias[o]) > 0 with w, x in bf16 (or tf32): it equals the
 * reference's fp32 convolution of the same bf16 values except where |logit| is within fp32 summation-order noise of zero;
 * for fp32 activations the tf32 products add a relative 2^-10 per term (tolerances in tests/test_gpu_head.py). */
int zp_head_decode(zp_ctx* ctx, const void* x, int c1, const void* x_skip, int c2, int dtype, int B, int S,
                   int mask_ch, int bit0_ch, int n_bits, int ignore_bit,
                   const double* bbox, const int32_t* obj_ids, int obj_default,
                   uint16_t* codes, float* corr, int cap, int32_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif
