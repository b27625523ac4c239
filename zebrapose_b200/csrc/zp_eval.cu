// The two steps either side of the pose path, batched on the device:
//   * zp_final_bbox   = padding_Bbox + get_final_Bbox (zebrapose/bop_dataset_pytorch.py:123-139, 162-194): the crop box
//                       the decode kernel consumes, computed from detection boxes without leaving the GPU;
//   * zp_pose_errors  = ADD and ADI (zebrapose/lib/pysixd/pose_error.py:297-336 through zebrapose/metric.py:8-18,
//                       evaluated per crop at zebrapose/test.py:465-483) of B pose pairs against the model vertices.
// ADD is float64 like the reference.  ADI replaces the reference's cKDTree by an exact brute-force nearest neighbour:
// every (ground-truth point, estimated point) pair is one FP32 squared distance on packed FFMA2/FADD2 pairs; the
// points are translated by -t_est in float64 before the float32 rounding so their magnitude is the object's radius
// (|error| of a distance <= ~2e-5 mm for a 150 mm object; the tolerance is written in tests/test_gpu_eval.py).
#include <algorithm>
#include <cmath>
#include "zp_common.cuh"

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 ev_fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 ev_mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 ev_sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 ev_dup(float x) {
    return (f32x2)__float_as_uint(x) | ((f32x2)__float_as_uint(x) << 32);
}
__device__ __forceinline__ float ev_lo(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float ev_hi(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }

// ---------------------------------------------------------------------------------------------------------------
// crop boxes
// ---------------------------------------------------------------------------------------------------------------
// Python's int() on a float truncates toward zero; all arithmetic below is the reference's float64 expression tree.
__device__ __forceinline__ double ev_trunc(double x) { return (double)__double2ll_rz(x); }

__global__ void zp_bbox_kernel(const double* __restrict__ in, int B, double pad_ratio, int method, double max_x,
                               double max_y, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double bx = in[4 * b], by = in[4 * b + 1], bw = in[4 * b + 2], bh = in[4 * b + 3];
    if (pad_ratio > 0) {   // padding_Bbox (bop_dataset_pytorch.py:123-139)
        double x1 = bx, x2 = bx + bw, y1 = by, y2 = by + bh;
        double cx = 0.5 * (x1 + x2), cy = 0.5 * (y1 + y2);
        double h = y2 - y1, w = x2 - x1;
        double pw = ev_trunc(w * pad_ratio), ph = ev_trunc(h * pad_ratio);
        bx = ev_trunc(cx - pw / 2); by = ev_trunc(cy - ph / 2); bw = pw; bh = ph;
    }
    // get_final_Bbox (bop_dataset_pytorch.py:162-194)
    double x1 = bx, x2 = bx + bw, y1 = by, y2 = by + bh;
    if (method == ZP_CROP_SQUARE_RESIZE || method == ZP_CROP_RESIZE_BY_WARP_AFFINE) {
        double cx = 0.5 * (x1 + x2), cy = 0.5 * (y1 + y2);
        if (bh > bw) { x1 = cx - bh / 2; x2 = cx + bh / 2; }
        else { y1 = cy - bw / 2; y2 = cy + bw / 2; }
        x1 = ev_trunc(x1); y1 = ev_trunc(y1); x2 = ev_trunc(x2); y2 = ev_trunc(y2);
        bx = x1; by = y1; bw = x2 - x1; bh = y2 - y1;
    } else if (method == ZP_CROP_RESIZE) {
        x1 = fmax(x1, 0.0); y1 = fmax(y1, 0.0); x2 = fmin(x2, max_x); y2 = fmin(y2, max_y);
        x1 = ev_trunc(x1); y1 = ev_trunc(y1); x2 = ev_trunc(x2); y2 = ev_trunc(y2);
        bx = x1; by = y1; bw = x2 - x1; bh = y2 - y1;
    }
    out[4 * b] = bx; out[4 * b + 1] = by; out[4 * b + 2] = bw; out[4 * b + 3] = bh;
}

// ---------------------------------------------------------------------------------------------------------------
// input crops: get_roi (crop_resize / crop_square_resize + cv2.resize INTER_LINEAR) + ToTensor + Normalize
// (bop_dataset_pytorch.py:36-89, 110-121, 334-347).  A CTA = 16 output rows, a thread = one column (its tap position and
// weights computed once; the rows' taps and the 3 x 256 ToTensor+Normalize values sit in shared memory); the zero-padded square canvas of
// crop_square_resize is never materialised (a tap outside the copied rectangle reads 0).  The resize is OpenCV's 8-bit
// fixed-point bilinear, restated instruction for instruction so that the crop is bit-identical to the reference's:
// float32 tap position from a float64 product, 11-bit coefficients rounded half-to-even, columns clamped with the
// fraction reset, rows clamped without, vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2, and the
// INTER_AREA switch for an exact 2x2 decimation.  HBM-bound on the output write (3 * sizeof(out) bytes per pixel).
// ---------------------------------------------------------------------------------------------------------------
struct CropArgs {
    const uint8_t* images; int n_img, H, W;
    const int32_t* img_ids; const double* boxes;
    int B, cs, method, out_bf16, channels_last;
    float mean[3], stdv[3];
    void* out; uint8_t* out_u8;
};

struct CropGeom {      // canvas (cw x ch) and the rectangle of it that holds image pixels
    int cw, ch, rx1, ry1, rx2, ry2, x1, y1;
};

__device__ __forceinline__ CropGeom zp_crop_geom(const double* bb, int method, int H, int W) {
    CropGeom g;
    const long long bx = __double2ll_rz(bb[0]), by = __double2ll_rz(bb[1]);
    if (method == ZP_CROP_SQUARE_RESIZE) {
        const long long bw = max(__double2ll_rz(bb[2]), 0ll), bh = max(__double2ll_rz(bb[3]), 0ll);
        double fx1 = (double)bx, fx2 = (double)(bx + bw), fy1 = (double)by, fy2 = (double)(by + bh);
        const double cx = 0.5 * (fx1 + fx2), cy = 0.5 * (fy1 + fy2);
        if (bh > bw) { fx1 = cx - (double)bh / 2; fx2 = cx + (double)bh / 2; }
        else { fy1 = cy - (double)bw / 2; fy2 = cy + (double)bw / 2; }
        long long x1 = __double2ll_rz(fx1), y1 = __double2ll_rz(fy1), x2 = __double2ll_rz(fx2), y2 = __double2ll_rz(fy2);
        const long long side = max(bh, bw);
        long long rx1 = max(-x1, 0ll); x1 = max(x1, 0ll);
        long long rx2 = rx1 + min((long long)W - x1, x2 - x1);
        long long ry1 = max(-y1, 0ll); y1 = max(y1, 0ll);
        long long ry2 = ry1 + min((long long)H - y1, y2 - y1);
        g.cw = g.ch = (int)side;
        g.rx1 = (int)min(rx1, side); g.ry1 = (int)min(ry1, side);
        g.rx2 = (int)min(max(rx2, rx1), side); g.ry2 = (int)min(max(ry2, ry1), side);     // numpy clips the slice to the canvas
        g.x1 = (int)x1; g.y1 = (int)y1;
    } else {
        const long long bw = __double2ll_rz(bb[2]), bh = __double2ll_rz(bb[3]);
        const long long x1 = max(0ll, bx), y1 = max(0ll, by);
        long long x2 = min((long long)W, bx + bw), y2 = min((long long)H, by + bh);
        // the reference slices img[y1:y2, x1:x2] (bop_dataset_pytorch.py:86): a NEGATIVE stop (box entirely above / left of
        // the image) counts from the far edge in numpy, so such a box crops almost the whole image -- reproduced, not fixed
        if (x2 < 0) x2 = max((long long)W + x2, 0ll);
        if (y2 < 0) y2 = max((long long)H + y2, 0ll);
        g.cw = (int)max(x2 - x1, 0ll); g.ch = (int)max(y2 - y1, 0ll);
        g.rx1 = 0; g.ry1 = 0; g.rx2 = g.cw; g.ry2 = g.ch;
        g.x1 = (int)x1; g.y1 = (int)y1;
    }
    return g;
}

// cv2's per-axis tap: position, neighbour and the two 11-bit weights (reset = columns: fraction zeroed at the borders)
__device__ __forceinline__ void zp_resize_tap(int d, int dn, int sn, bool reset, int& i0, int& i1, int& w0, int& w1) {
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dn, (double)sn));
    float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);      // no FMA contraction: cv2 rounds twice
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (reset) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= sn - 1) { f = 0.f; s = sn - 1; }
        i0 = s; i1 = min(s + 1, sn - 1);
    } else {
        i0 = min(max(s, 0), sn - 1); i1 = min(max(s + 1, 0), sn - 1);
    }
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

constexpr int CROP_ROWS = 16;          // output rows per CTA: a thread owns a column, its tap and weights are computed once

template <bool BF16, bool CL, bool U8>
__global__ void __launch_bounds__(256) zp_crop_kernel(CropArgs a) {
    __shared__ float s_lut[3][256];      // ToTensor + Normalize of every uint8 value, per channel (exact IEEE divisions, once)
    __shared__ int s_row[CROP_ROWS][4];  // byte offset of source row r0 / r1 (-1: outside the copied rectangle), b0, b1
    const int b = blockIdx.y, cs = a.cs, tid = threadIdx.x;
    const int oy0 = blockIdx.x * CROP_ROWS;
    const CropGeom g = zp_crop_geom(a.boxes + 4 * (size_t)b, a.method, a.H, a.W);
    const int id = a.img_ids ? a.img_ids[b] : 0;
    const uint8_t* img = a.images + (size_t)min(max(id, 0), a.n_img - 1) * a.H * a.W * 3;
    const bool empty = g.cw <= 0 || g.ch <= 0;
    const bool area2 = g.cw == 2 * cs && g.ch == 2 * cs;     // exact 2x2 decimation: cv2 uses INTER_AREA's rounded mean
#pragma unroll
    for (int c = 0; c < 3; c++)          // ToTensor: uint8 -> float32 / 255; Normalize: (x - mean) / std, IEEE float32 like torch
        s_lut[c][tid] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)tid, 255.f), a.mean[c]), a.stdv[c]);
    auto row_off = [&](int r) { return (r >= g.ry1 && r < g.ry2 && g.y1 + r - g.ry1 < a.H) ? (g.y1 + r - g.ry1) * a.W * 3 : -1; };
    auto col_off = [&](int c) { return (c >= g.rx1 && c < g.rx2 && g.x1 + c - g.rx1 < a.W) ? (g.x1 + c - g.rx1) * 3 : -1; };
    if (tid < CROP_ROWS && oy0 + tid < cs && !empty) {
        int r0, r1, b0 = 0, b1 = 0;
        if (area2) { r0 = 2 * (oy0 + tid); r1 = r0 + 1; }
        else zp_resize_tap(oy0 + tid, cs, g.ch, false, r0, r1, b0, b1);
        s_row[tid][0] = row_off(r0); s_row[tid][1] = row_off(r1); s_row[tid][2] = b0; s_row[tid][3] = b1;
    }
    __syncthreads();
    const size_t plane = (size_t)cs * cs;
    for (int ox = tid; ox < cs; ox += blockDim.x) {
        int x0 = 0, x1 = 0, a0 = 0, a1 = 0;
        if (!empty) {
            if (area2) { x0 = 2 * ox; x1 = x0 + 1; }
            else zp_resize_tap(ox, cs, g.cw, true, x0, x1, a0, a1);
        }
        const int c0 = empty ? -1 : col_off(x0), c1 = empty ? -1 : col_off(x1);
        for (int j = 0; j < CROP_ROWS && oy0 + j < cs; j++) {
            const int ro0 = s_row[j][0], ro1 = s_row[j][1], b0 = s_row[j][2], b1 = s_row[j][3];
            // a tap outside the copied rectangle reads the zero padding of the canvas
            const uint8_t* q00 = img + ro0 + c0; const bool v00 = !empty && (ro0 | c0) >= 0;
            const uint8_t* q01 = img + ro0 + c1; const bool v01 = !empty && (ro0 | c1) >= 0;
            const uint8_t* q10 = img + ro1 + c0; const bool v10 = !empty && (ro1 | c0) >= 0;
            const uint8_t* q11 = img + ro1 + c1; const bool v11 = !empty && (ro1 | c1) >= 0;
            int t00[3], t01[3], t10[3], t11[3], v[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                t00[c] = v00 ? q00[c] : 0; t01[c] = v01 ? q01[c] : 0; t10[c] = v10 ? q10[c] : 0; t11[c] = v11 ? q11[c] : 0;
            }
            if (area2) {
#pragma unroll
                for (int c = 0; c < 3; c++) v[c] = (t00[c] + t01[c] + t10[c] + t11[c] + 2) >> 2;
            } else {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int S0 = t00[c] * a0 + t01[c] * a1, S1 = t10[c] * a0 + t11[c] * a1;
                    v[c] = ((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2) & 0xff;
                }
            }
            const size_t p = (size_t)(oy0 + j) * cs + ox;
            if (U8) {
                uint8_t* o = a.out_u8 + ((size_t)b * plane + p) * 3;
                o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
            }
            if (a.out) {
                const float f0 = s_lut[0][v[0]], f1 = s_lut[1][v[1]], f2 = s_lut[2][v[2]];
                if (CL) {
                    const size_t i3 = ((size_t)b * plane + p) * 3;
                    if (BF16) { __nv_bfloat16* o = (__nv_bfloat16*)a.out + i3; o[0] = __float2bfloat16_rn(f0); o[1] = __float2bfloat16_rn(f1); o[2] = __float2bfloat16_rn(f2); }
                    else { float* o = (float*)a.out + i3; o[0] = f0; o[1] = f1; o[2] = f2; }
                } else {
                    const size_t i0 = (size_t)b * 3 * plane + p;
                    if (BF16) { __nv_bfloat16* o = (__nv_bfloat16*)a.out + i0; o[0] = __float2bfloat16_rn(f0); o[plane] = __float2bfloat16_rn(f1); o[2 * plane] = __float2bfloat16_rn(f2); }
                    else { float* o = (float*)a.out + i0; o[0] = f0; o[plane] = f1; o[2 * plane] = f2; }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// ADD / ADI
// ---------------------------------------------------------------------------------------------------------------
constexpr int EV_PREP_THREADS = 256;
constexpr int EV_THREADS = 128;          // ADI: threads per CTA
constexpr int EV_Q = 4;                  // query (ground-truth) points per thread, in registers
constexpr int EV_QTILE = EV_THREADS * EV_Q;
constexpr int EV_CHUNK = 2048;           // target (estimated) points staged in shared memory at a time (24 KB)
constexpr float EV_FAR = 1.0e18f;        // padding coordinate: squared distance 1e36 stays finite in float32
constexpr float EV_INIT = 3.0e38f;

struct ErrArgs {
    const double* est; const double* gt;          // [B,12] R row-major | t
    const int32_t* obj_ids; int obj_default;
    const double* const* model_ptrs; const int* model_V;
    int B, Vs;                                    // Vs = plane stride of E/G (multiple of 4)
    float* E; float* G;                           // [B][3][Vs]: R_e p  and  R_g p + (t_g - t_e), float32
    double* add_partial; int n_prep_blocks;       // [B][n_prep_blocks]
    uint32_t* qmin;                               // [B][Vs] float32 bit patterns of the smallest squared distance per query
    int nsplit, tlen;                             // target ranges per crop (tlen multiple of 4)
    double* add_out; double* adi_out;
};

__device__ __forceinline__ double ev_block_sum(double v, double* s_red) {
    // fixed-shape tree: warp shuffles, then warp 0 over the per-warp sums (deterministic for a given launch shape)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double r = 0;
    if (w == 0) {
        r = lane < nw ? s_red[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;            // valid on thread 0
}

// One thread per (crop, vertex): both rigid transforms in float64 (pose_error.py:308-309, misc.py:895-905), the ADD
// term ||p_est - p_gt||, and the float32 point sets of the ADI search.
__global__ void __launch_bounds__(EV_PREP_THREADS)
zp_err_prepare_kernel(ErrArgs a) {
    __shared__ double s_red[EV_PREP_THREADS / 32];
    const int b = blockIdx.y;
    const int obj = a.obj_ids ? a.obj_ids[b] : a.obj_default;
    const bool okobj = obj >= 0 && obj < ZP_MAX_OBJECTS && a.model_ptrs[obj] != nullptr;
    const int V = okobj ? a.model_V[obj] : 0;
    const int i = blockIdx.x * EV_PREP_THREADS + threadIdx.x;
    double term = 0;
    if (i < V) {
        const double* p = a.model_ptrs[obj] + 3 * (size_t)i;
        const double x = p[0], y = p[1], z = p[2];
        const double* e = a.est + 12 * (size_t)b;
        const double* g = a.gt + 12 * (size_t)b;
        double re[3], rg[3], d2 = 0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            re[r] = e[3 * r] * x + e[3 * r + 1] * y + e[3 * r + 2] * z;
            rg[r] = g[3 * r] * x + g[3 * r + 1] * y + g[3 * r + 2] * z;
            const double d = (re[r] + e[9 + r]) - (rg[r] + g[9 + r]);
            d2 += d * d;
        }
        term = sqrt(d2);
        if (a.E) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                a.E[((size_t)b * 3 + r) * a.Vs + i] = (float)re[r];
                a.G[((size_t)b * 3 + r) * a.Vs + i] = (float)(rg[r] + (g[9 + r] - e[9 + r]));
            }
        }
    } else if (i < a.Vs && a.E) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            a.E[((size_t)b * 3 + r) * a.Vs + i] = EV_FAR;
            a.G[((size_t)b * 3 + r) * a.Vs + i] = 0.f;
        }
    }
    const double s = ev_block_sum(term, s_red);
    if (threadIdx.x == 0) a.add_partial[(size_t)b * a.n_prep_blocks + blockIdx.x] = s;
}

// CTA = (tile of EV_QTILE ground-truth points, target range, crop).  Estimated points stream through shared memory in
// SoA quads (three LDS.128 broadcasts feed 4 targets x EV_Q queries); per target pair: 3 FADD2 + FMUL2 + 2 FFMA2, then
// the running minimum.  No tensor cores: a min-reduction over distances is not a contraction.
__global__ void __launch_bounds__(EV_THREADS)
zp_adi_kernel(ErrArgs a) {
    __shared__ __align__(16) float s_t[3][EV_CHUNK];
    const int b = blockIdx.z;
    const int obj = a.obj_ids ? a.obj_ids[b] : a.obj_default;
    const bool okobj = obj >= 0 && obj < ZP_MAX_OBJECTS && a.model_ptrs[obj] != nullptr;
    const int V = okobj ? a.model_V[obj] : 0;
    const int V4 = (V + 3) & ~3;
    const int q0 = blockIdx.x * EV_QTILE;
    const int t_begin = blockIdx.y * a.tlen, t_end = min(V4, t_begin + a.tlen);
    if (q0 >= V || t_begin >= t_end) return;
    const float* Eb = a.E + (size_t)b * 3 * a.Vs;
    const float* Gb = a.G + (size_t)b * 3 * a.Vs;
    f32x2 qx[EV_Q], qy[EV_Q], qz[EV_Q];
    float m[EV_Q];
#pragma unroll
    for (int k = 0; k < EV_Q; k++) {
        const int i = min(q0 + k * EV_THREADS + threadIdx.x, a.Vs - 1);     // consecutive lanes, consecutive points
        qx[k] = ev_dup(Gb[i]); qy[k] = ev_dup(Gb[a.Vs + i]); qz[k] = ev_dup(Gb[2 * (size_t)a.Vs + i]);
        m[k] = EV_INIT;
    }
    for (int c0 = t_begin; c0 < t_end; c0 += EV_CHUNK) {
        const int cn = min(EV_CHUNK, t_end - c0);               // multiple of 4
        __syncthreads();
#pragma unroll
        for (int pl = 0; pl < 3; pl++) {
            const float4* src = reinterpret_cast<const float4*>(Eb + (size_t)pl * a.Vs + c0);
            for (int j = threadIdx.x; j < (cn >> 2); j += EV_THREADS) reinterpret_cast<float4*>(s_t[pl])[j] = src[j];
        }
        __syncthreads();
        const ulonglong2* sx = reinterpret_cast<const ulonglong2*>(s_t[0]);
        const ulonglong2* sy = reinterpret_cast<const ulonglong2*>(s_t[1]);
        const ulonglong2* sz = reinterpret_cast<const ulonglong2*>(s_t[2]);
#pragma unroll 2
        for (int t = 0; t < (cn >> 2); t++) {
            const ulonglong2 X = sx[t], Y = sy[t], Z = sz[t];
#pragma unroll
            for (int k = 0; k < EV_Q; k++) {
                const f32x2 dx0 = ev_sub2(X.x, qx[k]), dx1 = ev_sub2(X.y, qx[k]);
                const f32x2 dy0 = ev_sub2(Y.x, qy[k]), dy1 = ev_sub2(Y.y, qy[k]);
                const f32x2 dz0 = ev_sub2(Z.x, qz[k]), dz1 = ev_sub2(Z.y, qz[k]);
                const f32x2 e0 = ev_fma2(dz0, dz0, ev_fma2(dy0, dy0, ev_mul2(dx0, dx0)));
                const f32x2 e1 = ev_fma2(dz1, dz1, ev_fma2(dy1, dy1, ev_mul2(dx1, dx1)));
                m[k] = fminf(fminf(m[k], ev_lo(e0)), ev_hi(e0));
                m[k] = fminf(fminf(m[k], ev_lo(e1)), ev_hi(e1));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < EV_Q; k++) {
        const int i = q0 + k * EV_THREADS + threadIdx.x;
        if (i < V) {
            uint32_t* dst = a.qmin + (size_t)b * a.Vs + i;
            if (a.nsplit > 1) atomicMin(dst, __float_as_uint(m[k]));     // non-negative floats order like their bit patterns
            else *dst = __float_as_uint(m[k]);
        }
    }
}

__global__ void __launch_bounds__(EV_PREP_THREADS)
zp_err_final_kernel(ErrArgs a) {
    __shared__ double s_red[EV_PREP_THREADS / 32];
    const int b = blockIdx.x;
    const int obj = a.obj_ids ? a.obj_ids[b] : a.obj_default;
    const bool okobj = obj >= 0 && obj < ZP_MAX_OBJECTS && a.model_ptrs[obj] != nullptr;
    const int V = okobj ? a.model_V[obj] : 0;
    const int nb = (V + EV_PREP_THREADS - 1) / EV_PREP_THREADS;
    double s = 0;
    for (int j = threadIdx.x; j < nb; j += EV_PREP_THREADS) s += a.add_partial[(size_t)b * a.n_prep_blocks + j];
    const double add_sum = ev_block_sum(s, s_red);
    __shared__ double s_add;
    if (threadIdx.x == 0) {
        s_add = V > 0 ? add_sum / V : nan("");
        if (a.add_out) a.add_out[b] = s_add;
    }
    if (!a.adi_out) return;
    double q = 0;
    for (int i = threadIdx.x; i < V; i += EV_PREP_THREADS) q += sqrt((double)__uint_as_float(a.qmin[(size_t)b * a.Vs + i]));
    const double adi_sum = ev_block_sum(q, s_red);       // contains the __syncthreads that publish s_add
    // a NaN pose makes every ADD term NaN; the float32 minimum would silently skip NaN distances, so mirror it here
    if (threadIdx.x == 0) a.adi_out[b] = (V > 0 && s_add == s_add) ? adi_sum / V : nan("");
}

static int ews_reserve(zp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ews_bytes) return 0;
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->ews) cudaFree(ctx->ews);
    ctx->ews = nullptr; ctx->ews_bytes = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    ZP_CUDA(ctx, cudaMalloc(&ctx->ews, want));
    ctx->ews_bytes = want;
    return 0;
}

static size_t ev_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" {

int zp_final_bbox(zp_ctx* ctx, const double* det_boxes, int B, double padding_ratio, int resize_method, double max_x,
                  double max_y, double* out_boxes, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (B < 0 || !det_boxes || !out_boxes) ZP_FAIL(ctx, -1, "zp_final_bbox: bad argument");
    if (resize_method < ZP_CROP_RESIZE || resize_method > ZP_CROP_KEEP)
        ZP_FAIL(ctx, -1, "zp_final_bbox: unknown resize method %d", resize_method);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    zp_bbox_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(det_boxes, B, padding_ratio, resize_method, max_x, max_y, out_boxes);
    ZP_CHECK_LAUNCH(ctx, "zp_bbox_kernel");
    return 0;
}

int zp_crop_input(zp_ctx* ctx, const uint8_t* images, int n_img, int H, int W, const int32_t* img_ids, const double* boxes,
                  int B, int crop_size, int resize_method, const float* mean3, const float* std3, int out_dtype,
                  int channels_last, void* out, uint8_t* out_u8, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (B < 0 || !images || !boxes || (!out && !out_u8) || n_img < 1 || H < 1 || W < 1) ZP_FAIL(ctx, -1, "zp_crop_input: bad argument");
    if (crop_size < 1 || crop_size > 4096) ZP_FAIL(ctx, -1, "zp_crop_input: bad crop size %d", crop_size);
    if ((long long)H * W * 3 > 0x7fffffffll) ZP_FAIL(ctx, -1, "zp_crop_input: image of %d x %d too large (32-bit byte offsets)", H, W);
    if (resize_method != ZP_CROP_RESIZE && resize_method != ZP_CROP_SQUARE_RESIZE)
        ZP_FAIL(ctx, -1, "zp_crop_input: resize method %d not supported (crop_resize | crop_square_resize; the warp-affine variant is not on this path)", resize_method);
    if (out_dtype != ZP_DTYPE_F32 && out_dtype != ZP_DTYPE_BF16) ZP_FAIL(ctx, -1, "zp_crop_input: bad output dtype %d", out_dtype);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    CropArgs a{};
    a.images = images; a.n_img = n_img; a.H = H; a.W = W; a.img_ids = img_ids; a.boxes = boxes;
    a.B = B; a.cs = crop_size; a.method = resize_method; a.out_bf16 = out_dtype == ZP_DTYPE_BF16; a.channels_last = channels_last != 0;
    const float dm[3] = {0.485f, 0.456f, 0.406f}, ds[3] = {0.229f, 0.224f, 0.225f};       // bop_dataset_pytorch.py:336
    for (int c = 0; c < 3; c++) { a.mean[c] = mean3 ? mean3[c] : dm[c]; a.stdv[c] = std3 ? std3[c] : ds[c]; }
    a.out = out; a.out_u8 = out_u8;
    ZP_TIME_BEGIN(ctx, (cudaStream_t)stream);
    const dim3 grid((crop_size + CROP_ROWS - 1) / CROP_ROWS, B);
    cudaStream_t st = (cudaStream_t)stream;
    const int variant = (a.out_bf16 ? 4 : 0) | (a.channels_last ? 2 : 0) | (out_u8 ? 1 : 0);
    switch (variant) {      // output format fixed at compile time: no per-pixel branches in the gather / blend loop
        case 0: zp_crop_kernel<false, false, false><<<grid, 256, 0, st>>>(a); break;
        case 1: zp_crop_kernel<false, false, true><<<grid, 256, 0, st>>>(a); break;
        case 2: zp_crop_kernel<false, true, false><<<grid, 256, 0, st>>>(a); break;
        case 3: zp_crop_kernel<false, true, true><<<grid, 256, 0, st>>>(a); break;
        case 4: zp_crop_kernel<true, false, false><<<grid, 256, 0, st>>>(a); break;
        case 5: zp_crop_kernel<true, false, true><<<grid, 256, 0, st>>>(a); break;
        case 6: zp_crop_kernel<true, true, false><<<grid, 256, 0, st>>>(a); break;
        default: zp_crop_kernel<true, true, true><<<grid, 256, 0, st>>>(a); break;
    }
    ZP_CHECK_LAUNCH(ctx, "zp_crop_kernel");
    return 0;
}

int zp_upload_model(zp_ctx* ctx, int obj_id, const double* pts_xyz, int V) {
    if (!ctx) return -1;
    if (obj_id < 0 || obj_id >= ZP_MAX_OBJECTS) ZP_FAIL(ctx, -1, "obj_id %d out of range [0,%d)", obj_id, ZP_MAX_OBJECTS);
    if (!pts_xyz || V < 1 || V > (1 << 22)) ZP_FAIL(ctx, -1, "zp_upload_model: bad vertex array (V = %d)", V);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    if (!ctx->d_model_ptrs) {
        ZP_CUDA(ctx, cudaMalloc((void**)&ctx->d_model_ptrs, ZP_MAX_OBJECTS * sizeof(double*)));
        ZP_CUDA(ctx, cudaMemset((void*)ctx->d_model_ptrs, 0, ZP_MAX_OBJECTS * sizeof(double*)));
        ZP_CUDA(ctx, cudaMalloc((void**)&ctx->d_model_V, ZP_MAX_OBJECTS * sizeof(int)));
        ZP_CUDA(ctx, cudaMemset(ctx->d_model_V, 0, ZP_MAX_OBJECTS * sizeof(int)));
    }
    ZpModel& m = ctx->models[obj_id];
    if (m.pts) cudaFree(m.pts);
    m.pts = nullptr; m.V = 0;
    ZP_CUDA(ctx, cudaMalloc((void**)&m.pts, (size_t)V * 3 * sizeof(double)));
    ZP_CUDA(ctx, cudaMemcpy(m.pts, pts_xyz, (size_t)V * 3 * sizeof(double), cudaMemcpyHostToDevice));
    m.V = V;
    ZP_CUDA(ctx, cudaMemcpy((void*)(ctx->d_model_ptrs + obj_id), &m.pts, sizeof(double*), cudaMemcpyHostToDevice));
    ZP_CUDA(ctx, cudaMemcpy(ctx->d_model_V + obj_id, &V, sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

int zp_pose_errors(zp_ctx* ctx, const double* poses_est, const double* poses_gt, const int32_t* obj_ids, int obj_default,
                   int B, double* add_out, double* adi_out, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (B < 0 || !poses_est || !poses_gt || (!add_out && !adi_out)) ZP_FAIL(ctx, -1, "zp_pose_errors: bad argument");
    int Vmax = 0;
    if (obj_ids) { for (const auto& m : ctx->models) Vmax = std::max(Vmax, m.V); }
    else {
        if (obj_default < 0 || obj_default >= ZP_MAX_OBJECTS || !ctx->models[obj_default].pts)
            ZP_FAIL(ctx, -1, "zp_pose_errors: no model uploaded for object slot %d", obj_default);
        Vmax = ctx->models[obj_default].V;
    }
    if (Vmax == 0) ZP_FAIL(ctx, -1, "zp_pose_errors: no model uploaded (zp_upload_model)");
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    ErrArgs a{};
    a.est = poses_est; a.gt = poses_gt; a.obj_ids = obj_ids; a.obj_default = obj_default;
    a.model_ptrs = ctx->d_model_ptrs; a.model_V = ctx->d_model_V;
    a.B = B; a.Vs = (Vmax + 3) & ~3;
    a.n_prep_blocks = (a.Vs + EV_PREP_THREADS - 1) / EV_PREP_THREADS;
    a.add_out = add_out; a.adi_out = adi_out;
    const size_t b_part = ev_align((size_t)B * a.n_prep_blocks * sizeof(double));
    const size_t b_pts = adi_out ? ev_align((size_t)B * 3 * a.Vs * sizeof(float)) : 0;
    const size_t b_min = adi_out ? ev_align((size_t)B * a.Vs * sizeof(uint32_t)) : 0;
    if (ews_reserve(ctx, b_part + 2 * b_pts + b_min)) return -2;
    char* p = (char*)ctx->ews;
    a.add_partial = (double*)p; p += b_part;
    if (adi_out) { a.E = (float*)p; p += b_pts; a.G = (float*)p; p += b_pts; a.qmin = (uint32_t*)p; }
    zp_err_prepare_kernel<<<dim3(a.n_prep_blocks, B), EV_PREP_THREADS, 0, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_err_prepare_kernel");
    if (adi_out) {
        const int qtiles = (Vmax + EV_QTILE - 1) / EV_QTILE;
        // enough CTAs for ~8 per SM: small batches split the target set and merge with atomicMin
        const long want = (long)ctx->sm_count * 8;
        int nsplit = (int)std::min<long>(std::max<long>(1, (want + (long)qtiles * B - 1) / ((long)qtiles * B)), std::max(1, a.Vs / 256));
        int tlen = ((a.Vs + nsplit - 1) / nsplit + 3) & ~3;
        nsplit = (a.Vs + tlen - 1) / tlen;
        a.nsplit = nsplit; a.tlen = tlen;
        if (nsplit > 1) ZP_CUDA(ctx, cudaMemsetAsync(a.qmin, 0x7f, (size_t)B * a.Vs * sizeof(uint32_t), st));   // 3.39e38f
        ZP_TIME_BEGIN(ctx, st);
        zp_adi_kernel<<<dim3(qtiles, nsplit, B), EV_THREADS, 0, st>>>(a);
        ZP_CHECK_LAUNCH(ctx, "zp_adi_kernel");
    }
    zp_err_final_kernel<<<B, EV_PREP_THREADS, 0, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_err_final_kernel");
    return 0;
}

}  // extern "C"
