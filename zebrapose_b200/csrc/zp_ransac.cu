// Kernel family 2: batched RANSAC-PnP (replaces cv2.solvePnPRansac(..., reprojectionError=2, iterationsCount=150,
// flags=SOLVEPNP_EPNP) + cv2.Rodrigues, /root/reference/zebrapose/binary_code_helper/CNN_output_to_pose.py:155-158).
//
//   zp_samples_kernel   one CTA per crop, one thread per hypothesis: cv::RNG(0xFFFFFFFFFFFFFFFF) replayed from a
//                       precomputed table of its raw 32-bit outputs; redraws on duplicates shift later hypotheses, which
//                       is resolved by a fixed-point iteration over a block prefix sum (exact; usually 1-2 rounds).
//                       Philox4x32-10 mode is counter based and needs no iteration.
//   zp_minimal_kernel   the FAST solver (zp_set_solver): a quad per hypothesis, float64 EPnP with a bisection / inverse-
//                       iteration null space.  The default solver is the exact replay of OpenCV's arithmetic, zp_cvsolve.cu.
//   zp_score_kernel     FP32-FMA bound: every correspondence x every hypothesis.  Correspondence tiles (SoA planes) are
//                       staged into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier) and kept in registers,
//                       hypotheses K[R|t] live in shared memory and are broadcast; the test is division free:
//                       (x - u z)^2 + (y - v z)^2 <= thr^2 z^2.  Counts: per-thread -> warp REDUX -> shared -> global.
//   zp_rs_*_kernel      cv2's loop state per crop (niters, maxGood, best) advanced wave by wave: "strictly greater" update +
//                       RANSACUpdateNumIters; near-ties decided on exact re-counts (cv2's arithmetic); crops that have reached
//                       their stopping iteration skip the later waves
//   zp_final_cl_kernel  the one-kernel final solve (zp_set_final_form 1 / 4, and the Gauss-Newton polish): EPnP on all inliers
//                       of the winner, a 4-CTA thread-block cluster per crop (or one CTA walking the same four partitions):
//                       inlier set (doubtful points by cv2's own arithmetic), reductions of the 52 EPnP sums through
//                       distributed shared memory, 16-lane null space, the three beta candidates on three lanes.  The
//                       default final solve is the split form, zp_finsplit.cu.
//
// Algorithmic FP32 work of scoring: 27 flop per (correspondence, hypothesis).
#include <climits>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include <cooperative_groups.h>
#include "zp_common.cuh"

// phase timestamps of thread 0 of CTA 0 (profiling aid, read with zp_debug_clocks): slots 0-9 final kernel, 10-15 minimal
// kernel, 16-19 inside the eigen-solver (last caller wins)
__device__ long long zp_dbg_clk[24];
#define ZP_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) zp_dbg_clk[i] = clock64(); } while (0)
#define ZP_EIG_STAMP(i) ZP_STAMP(i)
#include "zp_epnp.cuh"
#include "zp_proj.cuh"

// ---------------------------------------------------------------------------------------------------------------
// samples
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

constexpr int SMP_THREADS = 256;

constexpr int SMP_WINDOW = 2048;                 // raw RNG values staged in shared memory (150 x 5 draws need ~760)

// draws m distinct indices in [0,n) from the raw stream starting at table position `pos`; returns the number of raw
// values consumed, or -1 if the table would be overrun.  The first SMP_WINDOW values come from shared memory.
__device__ __forceinline__ int draw_from_table(const uint32_t* __restrict__ tab, const uint32_t* s_tab, int n_tab, int pos,
                                               int n, int m, int* idx) {
    int p = pos;
#pragma unroll
    for (int j = 0; j < 8; j++) {                     // m <= 8; static indices keep idx[] in registers
        if (j < m) {
            int v;
            bool dup;
            do {
                if (p >= n_tab) return -1;
                const uint32_t raw = p < SMP_WINDOW ? s_tab[p] : tab[p];
                p++;
                v = (int)(raw % (uint32_t)n);
                dup = false;
#pragma unroll
                for (int q = 0; q < 8; q++) dup |= q < j && idx[q] == v;
            } while (dup);
            idx[j] = v;
        }
    }
    return p - pos;
}

// grid = B, block = SMP_THREADS; thread h handles hypotheses h, h + SMP_THREADS, ...
__global__ void __launch_bounds__(SMP_THREADS)
zp_samples_kernel(const int32_t* __restrict__ counts, int cap, int B, int H, int m, int mode, uint64_t seed,
                  const uint32_t* __restrict__ rng_tab, int n_tab, int32_t* __restrict__ samples) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(counts[b], cap);
    int32_t* out = samples + (size_t)b * H * m;
    if (n < m) {
        for (int i = tid; i < H * m; i += SMP_THREADS) out[i] = -1;
        return;
    }
    if (mode != ZP_SAMPLER_CV2) {                  // counter based: (attempt, hypothesis, crop, slot) -> value
        for (int h = tid; h < H; h += SMP_THREADS) {
            int idx[8];
            for (int j = 0; j < m; j++) {
                int v;
                bool dup;
                uint32_t attempt = 0;
                do {
                    uint32_t c[4] = {attempt++, (uint32_t)h, (uint32_t)b, (uint32_t)j};
                    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
                    v = (int)(c[0] % (uint32_t)n);
                    dup = false;
                    for (int q = 0; q < j; q++) dup |= idx[q] == v;
                } while (dup);
                idx[j] = v;
                out[h * m + j] = v;
            }
        }
        return;
    }
    // cv2 replay.  off[h] = first raw value hypothesis h consumes = m*h + (redraws of all earlier hypotheses): a fixed
    // point of "draw from the current offsets -> block exclusive scan of the consumed counts", usually reached in 1-2
    // rounds (a redraw needs a duplicate among m indices in [0, n)).
    __shared__ uint32_t s_tab[SMP_WINDOW];
    __shared__ int s_off[ZP_MAX_HYPOTHESES];
    __shared__ int s_wsum[SMP_THREADS / 32];
    __shared__ int s_flag;
    for (int i = tid; i < SMP_WINDOW && i < n_tab; i += SMP_THREADS) s_tab[i] = rng_tab[i];
    for (int h = tid; h < H; h += SMP_THREADS) s_off[h] = m * h;
    if (tid == 0) s_flag = 0;
    __syncthreads();
    bool overflow = false;
    for (int round = 0; round <= H; round++) {
        int carry = 0, changed = 0, over = 0;
        for (int c0 = 0; c0 < H; c0 += SMP_THREADS) {         // H <= SMP_THREADS: one pass
            const int h = c0 + tid;
            int cons = 0;
            if (h < H) {
                int idx[8];
                cons = draw_from_table(rng_tab, s_tab, n_tab, s_off[h], n, m, idx);
                if (cons < 0) { cons = m; over = 1; }
            }
            int incl = cons;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int base = carry;
            for (int w = 0; w < warp; w++) base += s_wsum[w];
            const int off = base + incl - cons;
            if (h < H) { changed |= s_off[h] != off; s_off[h] = off; }
            for (int w = 0; w < SMP_THREADS / 32; w++) carry += s_wsum[w];
            __syncthreads();
        }
        if (changed | over) atomicOr(&s_flag, changed | (over << 1));
        __syncthreads();
        const int f = s_flag;
        __syncthreads();
        if (tid == 0) s_flag = 0;
        if (f & 2) { overflow = true; break; }
        if (!(f & 1)) break;
    }
    __syncthreads();
    if (!overflow) {
        for (int h = tid; h < H; h += SMP_THREADS) {
            int idx[8];
            draw_from_table(rng_tab, s_tab, n_tab, s_off[h], n, m, idx);
#pragma unroll
            for (int j = 0; j < 8; j++) if (j < m) out[h * m + j] = idx[j];
        }
        return;
    }
    if (tid == 0) {                                // pathological (tiny n): sequential generator, no table
        uint64_t state = 0xFFFFFFFFFFFFFFFFull;
        for (int h = 0; h < H; h++) {
            int idx[8];
            for (int j = 0; j < m; j++) {
                int v;
                bool dup;
                do {
                    state = (uint64_t)(uint32_t)state * 4164903690ull + (state >> 32);
                    v = (int)((uint32_t)state % (uint32_t)n);
                    dup = false;
                    for (int q = 0; q < j; q++) dup |= idx[q] == v;
                } while (dup);
                idx[j] = v;
                out[h * m + j] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// minimal solver
// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// projection matrices and the inlier predicate shared by the minimal solver (writes P), scoring and the final solve
// ---------------------------------------------------------------------------------------------------------------
// packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): two independent IEEE fma.rn per instruction, one issue slot and
// one 64-bit operand read per source
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 zp_fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 zp_mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 zp_sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 zp_pack2(float lo, float hi) {
    return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ bool zp_is_inlier(const float4& p0, const float4& p1, const float4& p2, float u, float v,
                                             float X, float Y, float Z) {
    return __float_as_int(zp_inlier_d(p0, p1, p2, u, v, X, Y, Z)) < 0;
}


constexpr int MIN_THREADS = 128;                 // 4 lanes per hypothesis -> 32 hypotheses per CTA
constexpr int ZP_MAX_M = 8;
// doubles of shared memory per hypothesis; 236 = 12 (mod 16) keeps the 8 quads of a warp on distinct banks (zp_epnp.cuh)
constexpr int MIN_HYP_DOUBLES = 236;
static_assert(MIN_HYP_DOUBLES >= ZP_SYM_DOUBLES + 24 + 48, "per-hypothesis shared memory layout");
constexpr int MIN_SMEM_BYTES = (MIN_THREADS / 4) * MIN_HYP_DOUBLES * (int)sizeof(double);

// One QUAD (4 lanes) per hypothesis.  The cheap serial setup (control points, 52 sums over the m points) is computed
// redundantly by the 4 lanes; the 12x12 null space is the quad-cooperative Jacobi; the three beta candidates run on
// lanes 0..2; the winner (smallest mean reprojection distance over the m points, EPnP's rule) writes the pose.
template <int MINB>
__global__ void __launch_bounds__(MIN_THREADS, MINB)
zp_minimal_kernel(const float* __restrict__ corr, int cap, const int32_t* __restrict__ counts,
                  const double* __restrict__ Kmat, const int32_t* __restrict__ samples, int B, int H, int h0, int hw,
                  const int32_t* __restrict__ crop_done /* nullable: crops that reached cv2's adaptive stop */,
                  const int32_t* __restrict__ rs /* nullable: [B,4] loop state, rs[4b] = the crop's current niters */, int m,
                  double inv_thr, double* __restrict__ hyp_poses, float* __restrict__ hyp_P,
                  int32_t* __restrict__ hyp_inliers /* nullable: zeroed here so the scoring launch needs no memset */) {
    extern __shared__ __align__(16) double s_min[];            // per hypothesis: z[12x13] | d[12] | e[12] | V[4x12]
    const int tid = threadIdx.x, lane = tid & 31, q = lane & 3, quad = tid >> 2;
    double* s_z = s_min + (size_t)quad * MIN_HYP_DOUBLES;
    double* s_Vq = s_z + ZP_SYM_DOUBLES + 24;
    // hypotheses [h0, h0 + hw) of every crop: local index -> (crop, hypothesis)
    const int total = B * hw;
    const int g_raw = blockIdx.x * (MIN_THREADS / 4) + quad;
    const int gl = g_raw < total ? g_raw : total - 1;   // dead quads shadow the last hypothesis (all lanes take part in shuffles)
    const int b = gl / hw;
    const int g = b * H + h0 + (gl - b * hw);
    const bool live = g_raw < total && !(crop_done && crop_done[b]) && !(rs && h0 + (gl - b * hw) >= rs[4 * b]);
    if (!__syncthreads_or(live)) return;           // every hypothesis of this CTA belongs to a finished crop
    ZP_STAMP(10);
    const int32_t* sidx = samples + (size_t)g * m;
    const int n = min(counts[b], cap);
    bool valid = n >= m;
    for (int j = 0; j < m; j++) valid = valid && sidx[j] >= 0 && sidx[j] < n;
    const double* Kb = Kmat + 9 * (size_t)b;
    const ZpCam cam{Kb[0], Kb[4], Kb[2], Kb[5]};
    const float* cb = corr + (size_t)b * 5 * cap;
    double X[ZP_MAX_M], Y[ZP_MAX_M], Z[ZP_MAX_M], U[ZP_MAX_M], Vv[ZP_MAX_M];
    double c0[3] = {0, 0, 0};
    for (int j = 0; j < m; j++) {
        int i = valid ? sidx[j] : 0;
        U[j] = valid ? (double)cb[i] : (double)j;
        Vv[j] = valid ? (double)cb[cap + i] : (double)(j * j);
        X[j] = valid ? (double)cb[2 * (size_t)cap + i] : (double)j;
        Y[j] = valid ? (double)cb[3 * (size_t)cap + i] : (double)(j & 1);
        Z[j] = valid ? (double)cb[4 * (size_t)cap + i] : (double)(j & 2);
        c0[0] += X[j]; c0[1] += Y[j]; c0[2] += Z[j];
    }
    c0[0] /= m; c0[1] /= m; c0[2] /= m;
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < m; j++) {
        double d[3] = {X[j] - c0[0], Y[j] - c0[1], Z[j] - c0[2]};
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) C[3 * r + c] = fma(d[r], d[c], C[3 * r + c]);
    }
    ZpControl cp;
    zp_control_points(c0, C, (double)m, cp);
    ZpHorn hs;
    double a_first[4];
    ZP_STAMP(11);
    {
        ZpSums sums;
        for (int e = 0; e < 10; e++) { sums.s0[e] = 0; sums.sx[e] = 0; sums.sy[e] = 0; sums.sr[e] = 0; }
        for (int e = 0; e < 12; e++) sums.w[e] = 0;
        sums.n = m;
        for (int j = 0; j < m; j++) {
            double a[4];
            zp_alphas(cp, X[j], Y[j], Z[j], a);
            if (j == 0) { a_first[0] = a[0]; a_first[1] = a[1]; a_first[2] = a[2]; a_first[3] = a[3]; }
            zp_accumulate(sums, a, cam.uc - U[j], cam.vc - Vv[j], X[j] - c0[0], Y[j] - c0[1], Z[j] - c0[2]);
        }
        zp_horn_inputs(sums, hs);
        // the 40 M^T M sums go through shared memory (the V slot is free until the solver returns) so that the fill can
        // index them dynamically; lane q of the quad stores the q-th group of ten
        {
            double* S = s_Vq;
#pragma unroll
            for (int e = 0; e < 10; e++) S[10 * q + e] = q == 0 ? sums.s0[e] : q == 1 ? sums.sx[e] : q == 2 ? sums.sy[e] : sums.sr[e];
            __syncwarp(0xFu << (lane & 28));
        }
        zp_nullspace4<4>(ZpSym12{s_z}, s_z + ZP_SYM_DOUBLES, s_z + ZP_SYM_DOUBLES + 12, s_Vq, cam, q, 0xFu << (lane & 28), s_Vq);
    }
    __syncwarp();
    ZP_STAMP(12);
    ZpMat V{s_Vq, 1};
    double L[60], rho[6];
    zp_L_rho(V, cp, L, rho);
    ZP_STAMP(13);
    double R[9], t[3];
    bool ok = zp_candidate(q < 3 ? q : 2, L, rho, V, hs, a_first, c0, R, t);
    double err = 0;
    for (int j = 0; j < m; j++) err += zp_reproj_dist(R, t, cam, X[j], Y[j], Z[j], U[j], Vv[j]);
    err /= m;
    ok = ok && err == err && q < 3;
    ZP_STAMP(14);
    // EPnP's choice: N = 1; if (err2 < err1) N = 2; if (err3 < err_N) N = 3 -- with non-finite candidates skipped
    const int base = lane & ~3;
    int pick = -1;
    double pe = 0;
    for (int c = 0; c < 3; c++) {
        double ec = __shfl_sync(0xffffffffu, err, base + c);
        int oc = __shfl_sync(0xffffffffu, (int)ok, base + c);
        if (oc && (pick < 0 || ec < pe)) { pick = c; pe = ec; }
    }
    if (live && q == 0 && hyp_inliers) hyp_inliers[g] = 0;
    if (live) {
        double* out = hyp_poses + (size_t)g * 12;
        float4* outP = (float4*)(hyp_P + (size_t)g * 24);       // every element twice: (P,P) pairs for FFMA2
        if (!valid || pick < 0) {
            if (q == 0) {
                for (int e = 0; e < 12; e++) out[e] = nan("");
                for (int e = 0; e < 6; e++) outP[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (q == pick) {
            double pose[12];
            for (int e = 0; e < 9; e++) pose[e] = R[e];
            for (int e = 0; e < 3; e++) pose[9 + e] = t[e];
            for (int e = 0; e < 12; e++) out[e] = pose[e];
            float P[12];
            zp_make_P(pose, Kb, inv_thr, P);
            for (int e = 0; e < 6; e++) outP[e] = make_float4(P[2 * e], P[2 * e], P[2 * e + 1], P[2 * e + 1]);
        }
    }
    ZP_STAMP(15);
}

// ---------------------------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------------------------
constexpr int SC_GROUP = 128;                      // threads that together cover one tile of correspondences
// SC_NG (template parameter) = warp-groups per CTA that split the hypotheses of one item: 1 for big batches; 4 when
// there are only a few items per CTA slot (64 crops = 832 items on 740 slots: with one group per CTA some CTAs get two
// items while the others idle after one -- 56 % balance; four groups per CTA make the CTAs 4x fewer and the items 4x
// shorter, so every CTA sees 4-5 items)
constexpr int SC_PPT = 8;                          // correspondences per thread (registers)
constexpr int SC_TILE = SC_GROUP * SC_PPT;         // correspondences per work item
constexpr int SC_HB = 160;                         // hypotheses staged in shared memory at a time (multiple of 32)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// float64 poses -> float32 projection matrices (only for zp_score with caller-supplied poses; the RANSAC chain gets P
// straight from the minimal solver)
__global__ void zp_poses_to_P_kernel(const double* __restrict__ poses, const double* __restrict__ K, int B, int H,
                                     double inv_thr, float* __restrict__ hyp_P) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * H) return;
    float P[12];
    zp_make_P(poses + (size_t)g * 12, K + 9 * (size_t)(g / H), inv_thr, P);
    float4* o = (float4*)(hyp_P + (size_t)g * 24);
    for (int e = 0; e < 6; e++) o[e] = make_float4(P[2 * e], P[2 * e], P[2 * e + 1], P[2 * e + 1]);
}

constexpr int SC_MAXCLS = 16;                      // hypothesis classes (chunks) a tile's work is cut into

struct ScoreArgs {
    const float* corr; int cap; const int32_t* counts; const float* hyp_P;
    int B, H; float inv_thr; int32_t* hyp_inliers; int* counters; int n_items;
    int h0, hw;                  // this launch scores hypotheses [h0, h0 + hw) ...
    const int32_t* crop_done;    // ... of the crops that have not reached cv2's adaptive stop (nullable: all)
    const int32_t* rs;           // nullable [B,4]: rs[4b] = the crop's current niters; hypotheses at or past it are not scored
    int n_cls;                   // a tile's hypotheses are cut into n_cls chunks [cls_off[c], cls_off[c+1]) (relative to h0,
    int cls_off[SC_MAXCLS + 1];  // each <= SC_HB); the queue hands out all tiles of chunk 0, then of chunk 1, ...
};

// Persistent CTAs pulling work items (hypothesis chunk c, crop b, tile of SC_TILE correspondences) from a global ticket
// counter.  The queue is ordered chunk-major; inside a chunk tile-major (w -> b = w % B, tile = w / B), so the empty tiles of
// short lists sit at the end of each segment; thread 0 draws tickets until it holds a live item (no CTA-wide barrier per
// skipped ticket).
// Per item: thread 0 issues TMA bulk copies of the 5 correspondence planes and of the chunk's projection matrices into
// shared memory (mbarrier completion); each thread keeps SC_PPT correspondences in registers and walks the hypotheses
// (3 x LDS.128 broadcast each); the sign bits of d are funnel-shifted into one register (1 instruction per evaluation),
// popc'ed, warp-reduced with REDUX and accumulated lane-distributed (lane h%32 owns hypothesis h).
template <int SC_NG>
__global__ void __launch_bounds__(SC_GROUP * SC_NG) zp_score_kernel(ScoreArgs a) {
    constexpr int SC_THREADS = SC_GROUP * SC_NG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_pts = (float*)smem_raw;                                   // [5][SC_TILE]
    ulonglong2* s_P = (ulonglong2*)(s_pts + 5 * SC_TILE);              // [SC_HB][6]: (P,P) pairs, 96 B per hypothesis
    int* s_cnt = (int*)(s_P + 6 * SC_HB);                              // [SC_HB]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_item[4];                                          // crop (-1: queue empty), first correspondence, hypotheses [begin, end)
    const int tid = threadIdx.x, lane = tid & 31, grp = tid / SC_GROUP, gt = tid % SC_GROUP;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int h = tid; h < SC_HB; h += SC_THREADS) s_cnt[h] = 0;
    uint32_t phase = 0;
    const int H = a.H;
    const int total = a.n_items * a.n_cls;
    for (;;) {
        __syncthreads();                                               // everybody is done with s_item / the buffers
        if (tid == 0) {
            int ib = -1, istart = 0, ih0 = 0, ih1 = 0, icnt = 0;
            for (;;) {
                int w = atomicAdd(&a.counters[0], 1);
                if (w >= total) break;
                const int cls = w / a.n_items;
                w -= cls * a.n_items;
                const int b = w % a.B, tile = w / a.B;
                if (a.crop_done && a.crop_done[b]) continue;
                const int n = min(a.counts[b], a.cap);
                if (tile * SC_TILE >= n) continue;
                const int hs = a.h0 + a.cls_off[cls];
                const int he = min(a.h0 + a.cls_off[cls + 1], a.rs ? a.rs[4 * b] : INT_MAX);
                if (hs >= he) continue;
                ib = b; istart = tile * SC_TILE; ih0 = hs; ih1 = he; icnt = min(SC_TILE, n - istart);
                break;
            }
            s_item[0] = ib; s_item[1] = icnt; s_item[2] = ih0; s_item[3] = ih1;
            if (ib >= 0) {
                const float* cb = a.corr + (size_t)ib * 5 * a.cap + istart;
                const uint32_t bytes = (uint32_t)((icnt + 3) & ~3) * 4u;   // 16-byte granules; cap % 4 == 0 keeps it in bounds
                const uint32_t pbytes = (uint32_t)(ih1 - ih0) * 96u;
                mbar_expect_tx(&s_bar, 5 * bytes + pbytes);
                for (int pl = 0; pl < 5; pl++) tma_load_1d(s_pts + pl * SC_TILE, cb + (size_t)pl * a.cap, bytes, &s_bar);
                tma_load_1d(s_P, a.hyp_P + ((size_t)ib * H + ih0) * 24, pbytes, &s_bar);
            }
        }
        __syncthreads();
        const int b = s_item[0];
        if (b < 0) break;
        const int cnt = s_item[1], h0 = s_item[2], hb = s_item[3] - s_item[2];
        f32x2 nu[SC_PPT / 2], nv[SC_PPT / 2], X[SC_PPT / 2], Y[SC_PPT / 2], Z[SC_PPT / 2];   // point pairs (j, j+1)
        mbar_wait(&s_bar, phase);
        phase ^= 1;
#pragma unroll
        for (int k = 0; k < SC_PPT / 2; k++) {
            float f[2][5];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                int i = gt + (2 * k + q) * SC_GROUP;
                bool live = i < cnt;
                // a dead slot gets u = 1e30: d >= +0 for every finite projection, never counted
                f[q][0] = live ? -(s_pts[i] * a.inv_thr) : -1e30f;
                f[q][1] = live ? -(s_pts[SC_TILE + i] * a.inv_thr) : 0.f;
                f[q][2] = live ? s_pts[2 * SC_TILE + i] : 0.f;
                f[q][3] = live ? s_pts[3 * SC_TILE + i] : 0.f;
                f[q][4] = live ? s_pts[4 * SC_TILE + i] : 0.f;
            }
            nu[k] = zp_pack2(f[0][0], f[1][0]); nv[k] = zp_pack2(f[0][1], f[1][1]);
            X[k] = zp_pack2(f[0][2], f[1][2]); Y[k] = zp_pack2(f[0][3], f[1][3]); Z[k] = zp_pack2(f[0][4], f[1][4]);
        }
        // group g takes hypotheses h = SC_NG * i + g; lane (i % 32) of every warp accumulates hypothesis i's count
        const int ni = (hb - grp + SC_NG - 1) / SC_NG;
        for (int iq = 0; iq < ni; iq += 32) {
            int acc = 0;
            const int iend = min(32, ni - iq);
#pragma unroll 2
            for (int il = 0; il < iend; il++) {
                const ulonglong2* pp = s_P + 6 * (SC_NG * (iq + il) + grp);
                const ulonglong2 q0 = pp[0], q1 = pp[1], q2 = pp[2], q3 = pp[3], q4 = pp[4], q5 = pp[5];
                // element-major order: every projection element is applied to all point pairs back to back, so
                // consecutive FFMA2s share a source operand (operand-reuse cache) -- an FFMA2 with three distinct
                // 64-bit register sources needs 3 even + 3 odd register reads and issues every 3 cycles instead of 2
                constexpr int NPAIR = SC_PPT / 2;
                f32x2 x[NPAIR], y[NPAIR], z[NPAIR];
#pragma unroll
                for (int k = 0; k < NPAIR; k++) z[k] = zp_fma2(q5.x, Z[k], q5.y);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) z[k] = zp_fma2(q4.y, Y[k], z[k]);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) z[k] = zp_fma2(q4.x, X[k], z[k]);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) x[k] = zp_fma2(q1.x, Z[k], q1.y);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) x[k] = zp_fma2(q0.y, Y[k], x[k]);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) x[k] = zp_fma2(q0.x, X[k], x[k]);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) y[k] = zp_fma2(q3.x, Z[k], q3.y);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) y[k] = zp_fma2(q2.y, Y[k], y[k]);
#pragma unroll
                for (int k = 0; k < NPAIR; k++) y[k] = zp_fma2(q2.x, X[k], y[k]);
                uint32_t bits = 0;
#pragma unroll
                for (int k = 0; k < NPAIR; k++) {
                    f32x2 dx = zp_fma2(nu[k], z[k], x[k]), dy = zp_fma2(nv[k], z[k], y[k]);
                    f32x2 e = zp_fma2(dx, dx, zp_mul2(dy, dy));
                    f32x2 d = zp_fma2(z[k] ^ 0x8000000080000000ull, z[k], e);   // e - z*z; the sign flip runs on the ALU pipe
                    bits = __funnelshift_l((uint32_t)d, bits, 1);
                    bits = __funnelshift_l((uint32_t)(d >> 32), bits, 1);
                }
                int c = __reduce_add_sync(0xffffffffu, __popc(bits));
                if (lane == il) acc += c;
            }
            if (acc) atomicAdd(&s_cnt[SC_NG * (iq + lane) + grp], acc);
        }
        __syncthreads();
        int32_t* out = a.hyp_inliers + (size_t)b * H + h0;
        for (int h = tid; h < hb; h += SC_THREADS) {
            int c = s_cnt[h];
            if (c) { atomicAdd(&out[h], c); s_cnt[h] = 0; }
        }
    }
    // the last CTA to leave re-arms the queue for the next launch
    if (tid == 0) {
        __threadfence();
        int done = atomicAdd(&a.counters[1], 1);
        if (done == (int)gridDim.x - 1) { a.counters[0] = 0; a.counters[1] = 0; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// winner selection + final solve on the inliers of the winner: one CTA per crop
// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// RANSAC control state: cv2's loop (RANSACPointSetRegistrator::run inside cv2.solvePnPRansac)
//     niters = H; for (it = 0; it < niters; it++) { good = count[it];
//         if (good > max(maxGood, m-1)) { best = it; maxGood = good; niters = RANSACUpdateNumIters(conf, (n-good)/n, m, niters); } }
// replayed wave by wave: hypotheses [h0, h1) have just been scored; the loop advances through them and stops where cv2
// would have stopped.  A crop is `done` once every remaining hypothesis lies at or past its niters -- the next waves
// skip it, and since cv2 never looks at those hypotheses the result does not depend on the wave plan.
// rs[b] = {niters, maxGood, best, iterations run}; one thread per crop (a handful of pow/log per crop).
// ---------------------------------------------------------------------------------------------------------------
__device__ inline int zp_update_iters(double p, double ep, int m, int maxit) {   // cv::RANSACUpdateNumIters
    p = fmin(fmax(p, 0.0), 1.0);
    ep = fmin(fmax(ep, 0.0), 1.0);
    double num = fmax(1.0 - p, ZP_DBL_MIN);
    double den = 1.0 - pow(1.0 - ep, (double)m);
    if (den < ZP_DBL_MIN) return 0;
    num = log(num);
    den = log(den);
    return (den >= 0 || -num >= maxit * (-den)) ? maxit : (int)rint(num / den);
}

__global__ void zp_rs_init_kernel(const int32_t* __restrict__ counts, int cap, int B, int H, int32_t* __restrict__ rs,
                                  int32_t* __restrict__ crop_done, int32_t* __restrict__ hyp_inliers_fill) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        const int n = min(counts[b], cap);
        rs[4 * b] = max(H, 1); rs[4 * b + 1] = 0; rs[4 * b + 2] = -1; rs[4 * b + 3] = 0;
        crop_done[b] = n < 6 ? 1 : 0;                 // CNN_output_to_pose.py:126: no RANSAC below six correspondences
        int32_t* tie = crop_done + B + 4 * (size_t)b;  // near-tie state of zp_rs_replay_kernel: nothing exact, nothing parked
        tie[0] = 0; tie[1] = -1; tie[2] = -1; tie[3] = 0;
    }
    if (hyp_inliers_fill)                              // hypotheses that are never run read back as -1
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * H; i += gridDim.x * blockDim.x) hyp_inliers_fill[i] = -1;
}

// One CTA per crop; every warp of it replays the same state machine on the same data (uniform control flow), so that the
// whole CTA can re-count a hypothesis when a decision needs exact counts.  Only "records" (counts above everything before
// them) can change the state, so a chunk of 32 counts is loaded coalesced, an inclusive prefix maximum flags the records, and
// the flagged ones are walked in order (pow / log only there).  (A first version -- one thread per crop walking the counts
// one by one -- took 54 us per wave.)
//
// Near-ties.  The scoring kernel's FP32 predicate may differ from cv2's float32 projectPoints error for a point within
// ~1e-4 px of the threshold (north_star allows 1e-3 px), so a count can be off by one, and when a hypothesis comes within
// RS_SLACK of the running maximum that is enough to flip cv2's strictly-greater decision (measured: 1 crop in 1536).  Such
// decisions are taken on EXACT counts instead: the CTA re-counts the hypothesis -- and, once, the current record holder --
// over all correspondences with the same predicate, re-deciding the doubtful points with cv2's own arithmetic
// (zp_inlier_exact), exactly as the final solve builds the winner's inlier set.  Clear decisions keep the FP32 counts.
constexpr int RS_THREADS = 256;                  // one CTA per crop: every warp replays the same state machine, all threads re-count
constexpr int RS_SLACK = 3;

// all RS_THREADS threads of the crop's CTA call this together (the replay's control flow is CTA-uniform); every thread gets the count
__device__ int zp_exact_count_cta(const float* __restrict__ cb, int cap, int n, const double* __restrict__ hp,
                                  const double* __restrict__ Kb, float inv_thr, float thr2, int* s_part) {
    const int tid = threadIdx.x;
    float P[12];
    zp_make_P(hp, Kb, (double)inv_thr, P);
    const float4 p0 = make_float4(P[0], P[1], P[2], P[3]), p1 = make_float4(P[4], P[5], P[6], P[7]),
                 p2 = make_float4(P[8], P[9], P[10], P[11]);
    const float *pu = cb, *pv = cb + cap, *pX = cb + 2 * (size_t)cap, *pY = cb + 3 * (size_t)cap, *pZ = cb + 4 * (size_t)cap;
    int c = 0;
    constexpr int PT = 4;                                  // points per thread and trip, their 20 loads up front (8: 170 registers, no faster)
    for (int i0 = 0; i0 < n; i0 += PT * RS_THREADS) {
        float fu[PT], fv[PT], fX[PT], fY[PT], fZ[PT];
#pragma unroll
        for (int q = 0; q < PT; q++) {
            const int i = i0 + RS_THREADS * q + tid;
            const bool ld = i < n;
            fu[q] = ld ? pu[i] : 0.f; fv[q] = ld ? pv[i] : 0.f; fX[q] = ld ? pX[i] : 0.f; fY[q] = ld ? pY[i] : 0.f; fZ[q] = ld ? pZ[i] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < PT; q++) {
            if (i0 + RS_THREADS * q + tid >= n) continue;
            const float d = zp_inlier_d(p0, p1, p2, fu[q] * inv_thr, fv[q] * inv_thr, fX[q], fY[q], fZ[q]);
            bool in = __float_as_int(d) < 0;
            const float z = fmaf(p2.x, fX[q], fmaf(p2.y, fY[q], fmaf(p2.z, fZ[q], p2.w)));
            if (fabsf(d) <= 1e-3f * z * z) in = zp_inlier_exact(hp, Kb[0], Kb[4], Kb[2], Kb[5], fu[q], fv[q], fX[q], fY[q], fZ[q], thr2);
            c += in;
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __syncthreads();                                       // the previous call's s_part has been read by everybody
    if ((tid & 31) == 0) s_part[tid >> 5] = c;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; w++) t += s_part[w];
    return t;
}

struct RsExact {            // inputs of the near-tie recount (corr == nullptr: FP32 counts decide everything)
    const float* corr; const double* K; const double* hyp_poses; float inv_thr, thr2;
    int park;               // 0: every near-tie is re-counted at once (test aid: the rule parking must reproduce)
    int32_t* tie;           // [B][4] per crop: {flags (1: maxGood is an exact count, 2: so is the parked state's), first parked iteration or -1, parked best, parked maxGood}
};

// cv2's loop state of one crop while the replay walks it (registers of the crop's warp, all lanes hold the same values)
struct RsState { int niters, maxgood, best, exact, last; };

// Walks iterations [from, to) of crop b.  Returns -1 when it reached `to` or cv2's stopping iteration (`stop` says which),
// or -- only with `park` -- the index of a near-tie that must be decided on exact counts while earlier near-ties are still
// parked: the caller resolves those first (an exact walk from the first parked one) and calls again from that index.
//
// Parking.  Before the first good hypothesis the counts are a few dozen and near-ties among them are frequent (measured: up to
// ~25 per crop, 8 us of re-counting each), but they are irrelevant as soon as a later hypothesis beats the running maximum
// by more than 2 RS_SLACK: it is a record on the exact counts too, whichever way the parked decisions went, and as long as the
// counts involved leave niters at the cap H the state after it is the same.  So a near-tie whose counts cannot move niters
// is decided on the FP32 counts and only remembered (`pend`: the state before the first such decision); a clear record
// forgets it; anything else that needs exact counts -- or the end of the last wave -- first replays from the parked state
// with re-counting.
__device__ int zp_rs_walk(int from, int to, RsState& s, bool& stop, bool park, int& pend_from, RsState& pend, const int32_t* hi,
                          int n, int m, int H, double conf, bool recount, const float* cb, int cap, const double* Kb,
                          const double* hp_crop, const RsExact& ex, int lane, int* s_part) {
    const int slack = recount ? RS_SLACK : 0;
    for (int c0 = from; c0 < to && !stop; c0 += 32) {
        const int h = c0 + lane;
        const int good = h < to ? hi[h] : INT_MIN;
        int v = good;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v = max(v, t);
        }
        int prev = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) prev = INT_MIN;
        // a hypothesis more than `slack` below an earlier count of its chunk cannot become a record on exact counts either
        unsigned bal = __ballot_sync(0xffffffffu, h < to && good > INT_MIN + slack && good + slack > max(max(prev, s.maxgood), m - 1));
        while (bal) {
            const int l = __ffs(bal) - 1;
            bal &= bal - 1;
            if (c0 + l >= s.niters) { stop = true; break; }       // cv2's loop has ended before this iteration
            int g = __shfl_sync(0xffffffffu, good, l);
            const int ref = max(s.maxgood, m - 1);
            if (g + slack <= ref) continue;                        // the maximum has moved on since the flags were taken
            int g_exact = 0;
            if (recount && g - ref <= slack) {                     // near-tie
                if (park && s.niters == H && zp_update_iters(conf, (double)(n - min(n, max(g, ref) + slack)) / n, m, H) == H) {
                    if (pend_from < 0) { pend_from = c0 + l; pend = s; }
                    if (g <= ref) continue;                        // FP32 decision for now; niters stays at the cap
                } else {
                    if (pend_from >= 0) return c0 + l;             // parked decisions come first
                    if (!s.exact && s.best >= 0) {
                        s.maxgood = zp_exact_count_cta(cb, cap, n, hp_crop + (size_t)s.best * 12, Kb, ex.inv_thr, ex.thr2, s_part);
                        s.exact = 1;
                    }
                    g = zp_exact_count_cta(cb, cap, n, hp_crop + (size_t)(c0 + l) * 12, Kb, ex.inv_thr, ex.thr2, s_part);
                    g_exact = 1;
                    if (g <= max(s.maxgood, m - 1)) continue;
                }
            } else if (pend_from >= 0 && g - ref > 2 * slack) {
                pend_from = -1;                                    // a record on any counts: the parked near-ties no longer matter
            }
            s.best = c0 + l; s.maxgood = g; s.last = c0 + l; s.exact = g_exact;
            s.niters = zp_update_iters(conf, (double)(n - g) / n, m, s.niters);
        }
        if (c0 + 32 >= s.niters) stop = true;
    }
    return -1;
}

__global__ void __launch_bounds__(RS_THREADS)
zp_rs_replay_kernel(const int32_t* __restrict__ counts, int cap, const int32_t* __restrict__ hyp_inliers, int B, int H, int h0,
                    int h1, int m, double conf, int select_mode, int32_t* __restrict__ rs, int32_t* __restrict__ crop_done, RsExact ex) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    __shared__ int s_part[RS_THREADS / 32];
    if (crop_done[b]) return;
    const int n = min(counts[b], cap);
    const int32_t* hi = hyp_inliers + (size_t)b * H;
    if (select_mode != ZP_SELECT_CV2_REPLAY) {             // most inliers, lowest index on ties, every hypothesis consulted
        if (threadIdx.x >= 32) return;
        int maxgood = rs[4 * b + 1], best = rs[4 * b + 2];
        for (int c0 = h0; c0 < h1; c0 += 32) {
            const int h = c0 + lane;
            const int good = h < h1 ? hi[h] : INT_MIN;
            int v = good;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v = max(v, t);
            }
            int prev = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == 0) prev = INT_MIN;
            const unsigned bal = __ballot_sync(0xffffffffu, h < h1 && good > max(max(prev, maxgood), m - 1));
            if (bal) {
                const int l = 31 - __clz(bal);
                best = c0 + l; maxgood = __shfl_sync(0xffffffffu, good, l);
            }
        }
        if (lane == 0) {
            rs[4 * b + 1] = maxgood; rs[4 * b + 2] = best; rs[4 * b + 3] = h1;
            if (h1 >= H) crop_done[b] = 1;
        }
        return;
    }
    const bool recount = ex.corr != nullptr;
    const float* cb = recount ? ex.corr + (size_t)b * 5 * cap : nullptr;
    const double* Kb = recount ? ex.K + 9 * (size_t)b : nullptr;
    const double* hp_crop = recount ? ex.hyp_poses + (size_t)b * H * 12 : nullptr;
    int32_t* tie = ex.tie + 4 * (size_t)b;
    RsState s, pend;
    s.niters = rs[4 * b]; s.maxgood = rs[4 * b + 1]; s.best = rs[4 * b + 2]; s.exact = tie[0] & 1; s.last = h0 - 1;
    int pend_from = tie[1];
    pend.niters = H; pend.best = tie[2]; pend.maxgood = tie[3]; pend.exact = (tie[0] >> 1) & 1; pend.last = h0 - 1;
    bool stop = false;
    int cur = h0;
    for (;;) {
        const int at = zp_rs_walk(cur, h1, s, stop, ex.park != 0, pend_from, pend, hi, n, m, H, conf, recount, cb, cap, Kb, hp_crop, ex, lane, s_part);
        // parked near-ties must be settled before a decision that needs exact counts, and before this crop's last wave ends
        const bool closing = at < 0 && (stop || h1 >= H || h1 >= s.niters);
        if (pend_from < 0 || (at < 0 && !closing)) break;
        const int upto = at >= 0 ? at : h1;
        s = pend;
        stop = false;
        int none = -1;
        RsState dummy = s;
        zp_rs_walk(pend_from, upto, s, stop, false, none, dummy, hi, n, m, H, conf, recount, cb, cap, Kb, hp_crop, ex, lane, s_part);
        pend_from = -1;
        if (at < 0 || stop) break;
        cur = at;
    }
    __syncthreads();                                       // every warp has read the crop's state before it is overwritten
    if (threadIdx.x == 0) {
        const int it = max(s.last + 1, max(h0, min(h1, s.niters)));
        rs[4 * b] = s.niters; rs[4 * b + 1] = s.maxgood; rs[4 * b + 2] = s.best; rs[4 * b + 3] = it;
        tie[0] = s.exact | (pend.exact << 1); tie[1] = pend_from; tie[2] = pend.best; tie[3] = pend.maxgood;
        if (h1 >= H || h1 >= s.niters) crop_done[b] = 1;
    }
}


template <int NV, int FIN_THREADS>
__device__ __forceinline__ void block_reduce(double* v, double* s_red /* [FIN_THREADS/32][NV] */, double* s_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double x = v[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) s_red[warp * NV + q] = x;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < NV; q += FIN_THREADS) {
        double t = 0;
        for (int w = 0; w < FIN_THREADS / 32; w++) t += s_red[w * NV + q];
        s_out[q] = t;
    }
    __syncthreads();
}


// ---------------------------------------------------------------------------------------------------------------
// Winner's inliers -> final EPnP, spread over a THREAD-BLOCK CLUSTER.  Round 1's form -- one CTA per crop -- kept 64 of the
// 148 SMs busy at BASELINE's 64-crop batches, and its point passes (inlier set, scatter, 52 EPnP sums, candidate errors:
// 83 of CTA 0's 127 us, profiles/r2k_final_phases.txt) ran on 4 warps that can each issue one FP64 instruction every ~3.3
// cycles: 172 us at 64 crops against 119 us for the cluster (693 vs 1280 us at 1024 crops, where the GPU is full anyway).  Here a
// crop's points are cut into FIN_VR = 4 "virtual ranks" of 4 warps each (16 warp chunks); with CL = 4 a virtual rank is a
// CTA of a 4-CTA cluster and the partial sums are combined through distributed shared memory, with CL = 1 one CTA walks the
// four virtual ranks in turn.  Partition, per-rank reduction tree and the order ((p0 + p1) + p2) + p3 are the same in
// both, so a crop's pose does not depend on which form ran (small batches take the cluster, saturating ones the single
// CTA).  The serial solver phases run redundantly in every CTA of a cluster -- same inputs, same bits, no broadcast.
// ---------------------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int FIN_VR = 4;              // virtual ranks per crop
constexpr int FCL_THREADS = 128;
constexpr int FCL_WARPS = FCL_THREADS / 32;

template <int CL>
__device__ __forceinline__ void fcl_sync() {
    if (CL > 1) cg::this_cluster().sync();
    else __syncthreads();
}

// s_part[vr_local][q] of every virtual rank -> s_sum[q] = ((p0 + p1) + p2) + p3, identical in every CTA
template <int CL, int NV>
__device__ __forceinline__ void fcl_total(double (*s_part)[56], double* s_sum) {
    constexpr int NR = FIN_VR / CL;
    fcl_sync<CL>();
    for (int q = threadIdx.x; q < NV; q += FCL_THREADS) {
        double t = 0;
#pragma unroll
        for (int vr = 0; vr < FIN_VR; vr++) {
            const double* src = &s_part[vr % NR][q];
            if (CL > 1) src = cg::this_cluster().map_shared_rank(src, vr / NR);
            t += *src;
        }
        s_sum[q] = t;
    }
    fcl_sync<CL>();
}

template <int CL>
__global__ void __launch_bounds__(FCL_THREADS) zp_final_cl_kernel(FinalArgs a) {
    constexpr int NR = FIN_VR / CL;                 // virtual ranks walked by this CTA
    const int b = blockIdx.x / CL, rank = blockIdx.x % CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_red[FCL_WARPS * 52];
    __shared__ double s_part[NR][56];               // per virtual rank: up to 52 sums | inlier count | first inlier index
    __shared__ double s_sum[56];
    __shared__ double s_V[48];
    __shared__ __align__(16) double s_eig[ZP_SYM_DOUBLES + 24];
    __shared__ ZpControl s_cp;
    __shared__ ZpSums s_sums;
    __shared__ double s_candR[3][9], s_candt[3][3];
    __shared__ int s_candok[3];
    __shared__ double s_pose[12];
    __shared__ int s_best, s_status;
    __shared__ int s_wcnt[NR][FCL_WARPS], s_wpre[NR][FCL_WARPS], s_nvr[NR];
    extern __shared__ __align__(8) unsigned char s_dynb[];
    uint16_t* s_idx = (uint16_t*)s_dynb;            // packed inlier indices: [NR][FCL_WARPS][chunk]

    const bool writer = rank == 0;
    double* out = a.poses + 12 * (size_t)b;
    const int n_raw = a.counts[b];
    const int n = min(n_raw, a.cap);
    if (tid == 0) {
        int best = -1, st = ZP_OK;
        if (n_raw == 0) st = ZP_NO_MASK_PIXELS;
        else if (n < 6) st = ZP_TOO_FEW_POINTS;       // CNN_output_to_pose.py:126
        else {
            best = a.rs[4 * b + 2];
            if (best < 0) st = ZP_RANSAC_NO_MODEL;
        }
        s_best = best; s_status = st;
        if (writer) {
            a.status[b] = st;
            if (a.best_idx) a.best_idx[b] = best;
            if (a.iters_run) a.iters_run[b] = n < 6 ? 0 : a.rs[4 * b + 3];
        }
    }
    __syncthreads();
    const int best = s_best;
    const int chunk = ((n + FIN_VR * FCL_THREADS - 1) / (FIN_VR * FCL_THREADS)) * 32;   // points per virtual warp
    if (best < 0) {     // no model: cv2 leaves rvec = tvec = 0 and the reference reports R = I, t = 0
        if (writer) {
            if (tid < 12) out[tid] = (tid == 0 || tid == 4 || tid == 8) ? 1.0 : 0.0;
            if (tid == 0) a.n_inliers[b] = 0;
            if (a.records && tid < 14)
                a.records[14 * (size_t)b + tid] = tid < 12 ? ((tid == 0 || tid == 4 || tid == 8) ? 1.0 : 0.0) : tid == 12 ? 0.0 : (double)s_status;
        }
        if (a.inlier_mask)
            for (int i = rank * FCL_THREADS + tid; i < a.cap; i += CL * FCL_THREADS) a.inlier_mask[(size_t)b * a.cap + i] = 0;
        return;
    }
    const float* cb = a.corr + (size_t)b * 5 * a.cap;
    const float *pu = cb, *pv = cb + a.cap, *pX = cb + 2 * (size_t)a.cap, *pY = cb + 3 * (size_t)a.cap, *pZ = cb + 4 * (size_t)a.cap;
    const double* Kb = a.K + 9 * (size_t)b;
    const double* hp = a.hyp_poses + ((size_t)b * a.H + best) * 12;
    float P[12];
    zp_make_P(hp, Kb, (double)a.inv_thr, P);
    const float4 p0 = make_float4(P[0], P[1], P[2], P[3]), p1 = make_float4(P[4], P[5], P[6], P[7]),
                 p2 = make_float4(P[8], P[9], P[10], P[11]);
    const float thr2 = a.thr2;
    double acc[52];
    // ---- pass 0: inlier set of the winner (same predicate as zp_score_kernel, doubtful points by cv2's arithmetic), centroid
    for (int vl = 0; vl < NR; vl++) {
        const int vw = (rank * NR + vl) * FCL_WARPS + warp;          // virtual warp: points [vw * chunk, (vw + 1) * chunk)
        uint16_t* seg = s_idx + ((size_t)vl * FCL_WARPS + warp) * chunk;
        acc[0] = acc[1] = acc[2] = 0;
        int my_first = 0x7fffffff, wcount = 0;
        for (int j0 = 0; j0 < chunk; j0 += 128) {                    // four 32-point groups per trip, their 20 loads up front
            float fu[4], fv[4], fX[4], fY[4], fZ[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = vw * chunk + j0 + 32 * q + lane;
                const bool ld = j0 + 32 * q < chunk && i < n;
                fu[q] = ld ? pu[i] : 0.f; fv[q] = ld ? pv[i] : 0.f; fX[q] = ld ? pX[i] : 0.f; fY[q] = ld ? pY[i] : 0.f; fZ[q] = ld ? pZ[i] : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (j0 + 32 * q >= chunk) break;                     // warp-uniform
                const int i = vw * chunk + j0 + 32 * q + lane;
                bool in = false;
                if (i < n) {
                    const float d = zp_inlier_d(p0, p1, p2, fu[q] * a.inv_thr, fv[q] * a.inv_thr, fX[q], fY[q], fZ[q]);
                    in = __float_as_int(d) < 0;
                    const float z = fmaf(p2.x, fX[q], fmaf(p2.y, fY[q], fmaf(p2.z, fZ[q], p2.w)));
                    if (fabsf(d) <= 1e-3f * z * z)                   // within ~1e-3 px of the threshold: cv2's own arithmetic decides
                        in = zp_inlier_exact(hp, Kb[0], Kb[4], Kb[2], Kb[5], fu[q], fv[q], fX[q], fY[q], fZ[q], thr2);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (a.inlier_mask && i < n) a.inlier_mask[(size_t)b * a.cap + i] = in;
                if (in) {
                    seg[wcount + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
                    my_first = min(my_first, i);
                    acc[0] += (double)fX[q]; acc[1] += (double)fY[q]; acc[2] += (double)fZ[q];
                }
                wcount += __popc(bal);
            }
        }
        my_first = __reduce_min_sync(0xffffffffu, my_first);
        if (lane == 0) { s_wcnt[vl][warp] = wcount; s_red[FCL_WARPS * 3 + warp] = (double)my_first; }
        block_reduce<3, FCL_THREADS>(acc, s_red, s_part[vl]);
        if (tid == 0) {
            int run = 0;
            double fi = 2147483647.0;
            for (int w = 0; w < FCL_WARPS; w++) { s_wpre[vl][w] = run; run += s_wcnt[vl][w]; fi = fmin(fi, s_red[FCL_WARPS * 3 + w]); }
            s_nvr[vl] = run;
            s_part[vl][3] = (double)run;
            s_part[vl][4] = fi;
        }
        __syncthreads();
    }
    if (a.inlier_mask)
        for (int i = n + rank * FCL_THREADS + tid; i < a.cap; i += CL * FCL_THREADS) a.inlier_mask[(size_t)b * a.cap + i] = 0;
    // totals: centroid sums and the inlier count add up, the first inlier index is a minimum
    {
        fcl_sync<CL>();
        if (tid < 5) {
            double t = tid == 4 ? 2147483647.0 : 0.0;
#pragma unroll
            for (int vr = 0; vr < FIN_VR; vr++) {
                const double* src = &s_part[vr % NR][tid];
                if (CL > 1) src = cg::this_cluster().map_shared_rank(src, vr / NR);
                t = tid == 4 ? fmin(t, *src) : t + *src;
            }
            s_sum[tid] = t;
        }
        fcl_sync<CL>();
    }
    const int ni = (int)s_sum[3];
    const int first = (int)s_sum[4];
    if (writer && tid == 0) a.n_inliers[b] = ni;
    if (ni < 4) {       // cannot happen after selection (good > m-1 >= 3) but keep the output defined
        if (writer) {
            if (tid < 12) out[tid] = hp[tid];
            if (a.records && tid < 14) a.records[14 * (size_t)b + tid] = tid < 12 ? hp[tid] : tid == 12 ? (double)ni : (double)s_status;
        }
        return;
    }
    const double c0[3] = {s_sum[0] / ni, s_sum[1] / ni, s_sum[2] / ni};
    // k-th inlier of local virtual rank vl (k < s_nvr[vl]) in its packed per-warp segments
    auto inlier_at = [&](int vl, int k) -> int {
        int w = 0;
#pragma unroll
        for (int q = 1; q < FCL_WARPS; q++) w += k >= s_wpre[vl][q];
        return s_idx[((size_t)vl * FCL_WARPS + w) * chunk + (k - s_wpre[vl][w])];
    };
    // ---- pass 1: scatter matrix
    for (int vl = 0; vl < NR; vl++) {
        for (int q = 0; q < 9; q++) acc[q] = 0;
        const int nv = s_nvr[vl];
#pragma unroll 4
        for (int k = tid; k < nv; k += FCL_THREADS) {
            const int i = inlier_at(vl, k);
            double d0 = pX[i] - c0[0], d1 = pY[i] - c0[1], d2 = pZ[i] - c0[2];
            acc[0] = fma(d0, d0, acc[0]); acc[1] = fma(d0, d1, acc[1]); acc[2] = fma(d0, d2, acc[2]);
            acc[4] = fma(d1, d1, acc[4]); acc[5] = fma(d1, d2, acc[5]); acc[8] = fma(d2, d2, acc[8]);
        }
        acc[3] = acc[1]; acc[6] = acc[2]; acc[7] = acc[5];
        block_reduce<9, FCL_THREADS>(acc, s_red, s_part[vl]);
    }
    fcl_total<CL, 9>(s_part, s_sum);
    if (tid == 0) {
        double C[9];
        for (int q = 0; q < 9; q++) C[q] = s_sum[q];
        zp_control_points(c0, C, (double)ni, s_cp);
    }
    __syncthreads();
    // ---- pass 2: the 52 EPnP sums
    const ZpCam cam{Kb[0], Kb[4], Kb[2], Kb[5]};
    for (int vl = 0; vl < NR; vl++) {
        ZpSums s;
        for (int q = 0; q < 10; q++) { s.s0[q] = 0; s.sx[q] = 0; s.sy[q] = 0; s.sr[q] = 0; }
        for (int q = 0; q < 12; q++) s.w[q] = 0;
        const ZpControl cp = s_cp;
        const int nv = s_nvr[vl];
        // the five loads of the next point are issued before the ~80 FP64 operations of the current one
        int k = tid;
        bool in = k < nv;
        int i = in ? inlier_at(vl, k) : 0;
        float fX = in ? pX[i] : 0.f, fY = in ? pY[i] : 0.f, fZ = in ? pZ[i] : 0.f, fu = in ? pu[i] : 0.f, fv = in ? pv[i] : 0.f;
        while (in) {
            const int k2 = k + FCL_THREADS;
            const bool in2 = k2 < nv;
            const int i2 = in2 ? inlier_at(vl, k2) : 0;
            const float gX = in2 ? pX[i2] : 0.f, gY = in2 ? pY[i2] : 0.f, gZ = in2 ? pZ[i2] : 0.f, gu = in2 ? pu[i2] : 0.f, gv = in2 ? pv[i2] : 0.f;
            {
                double X = fX, Y = fY, Z = fZ;
                double al[4];
                zp_alphas(cp, X, Y, Z, al);
                zp_accumulate(s, al, cam.uc - (double)fu, cam.vc - (double)fv, X - c0[0], Y - c0[1], Z - c0[2]);
            }
            k = k2; in = in2; fX = gX; fY = gY; fZ = gZ; fu = gu; fv = gv;
        }
        for (int q = 0; q < 10; q++) { acc[q] = s.s0[q]; acc[10 + q] = s.sx[q]; acc[20 + q] = s.sy[q]; acc[30 + q] = s.sr[q]; }
        for (int q = 0; q < 12; q++) acc[40 + q] = s.w[q];
        block_reduce<52, FCL_THREADS>(acc, s_red, s_part[vl]);
    }
    fcl_total<CL, 52>(s_part, s_sum);
    if (tid < 10) { s_sums.s0[tid] = s_sum[tid]; s_sums.sx[tid] = s_sum[10 + tid]; s_sums.sy[tid] = s_sum[20 + tid]; s_sums.sr[tid] = s_sum[30 + tid]; }
    if (tid < 12) s_sums.w[tid] = s_sum[40 + tid];
    if (tid == 0) s_sums.n = ni;
    __syncthreads();
    // ---- 12x12 null space on 16 lanes of warp 0, the three beta candidates on three lanes (every CTA of a cluster: same bits)
    if (tid < 32) {
        if (tid < 16) zp_nullspace4<16>(ZpSym12{s_eig}, s_eig + ZP_SYM_DOUBLES, s_eig + ZP_SYM_DOUBLES + 12, s_sums.s0, cam, tid, 0xFFFFu, s_V);
        __syncwarp();
        if (lane < 3) {
            ZpMat V{s_V, 1};
            double L[60], rho[6], af[4];
            zp_L_rho(V, s_cp, L, rho);
            ZpHorn hs;
            zp_horn_inputs(s_sums, hs);
            zp_alphas(s_cp, pX[first], pY[first], pZ[first], af);
            s_candok[lane] = zp_candidate(lane, L, rho, V, hs, af, c0, s_candR[lane], s_candt[lane]) ? 1 : 0;
        }
    }
    __syncthreads();
    // ---- pass 3: mean reprojection distance of the three candidates, pick the best
    {
        double cR[3][9], ct[3][3];
        bool cok[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            cok[c] = s_candok[c] != 0;
#pragma unroll
            for (int e = 0; e < 9; e++) cR[c][e] = cok[c] ? s_candR[c][e] : (e % 4 == 0 ? 1.0 : 0.0);
#pragma unroll
            for (int e = 0; e < 3; e++) ct[c][e] = cok[c] ? s_candt[c][e] : (e == 2 ? 1.0 : 0.0);
        }
        for (int vl = 0; vl < NR; vl++) {
            for (int q = 0; q < 3; q++) acc[q] = 0;
            const int nv = s_nvr[vl];
#pragma unroll 2
            for (int k = tid; k < nv; k += FCL_THREADS) {
                const int i = inlier_at(vl, k);
                const double X = pX[i], Y = pY[i], Z = pZ[i], u = pu[i], v = pv[i];
#pragma unroll
                for (int c = 0; c < 3; c++) acc[c] += zp_reproj_dist(cR[c], ct[c], cam, X, Y, Z, u, v);
            }
            block_reduce<3, FCL_THREADS>(acc, s_red, s_part[vl]);
        }
    }
    fcl_total<CL, 3>(s_part, s_sum);
    if (tid == 0) {
        int pick = -1;
        double be = 0;
        for (int c = 0; c < 3; c++) {
            if (!s_candok[c]) continue;
            double e = s_sum[c] / ni;
            if (!(e == e)) continue;
            if (pick < 0 || e < be) { pick = c; be = e; }
        }
        if (pick < 0) for (int e = 0; e < 12; e++) s_pose[e] = hp[e];
        else {
            for (int e = 0; e < 9; e++) s_pose[e] = s_candR[pick][e];
            for (int e = 0; e < 3; e++) s_pose[9 + e] = s_candt[pick][e];
        }
    }
    __syncthreads();
    // ---- optional Gauss-Newton polish of the pixel reprojection error over the inliers (north_star extension)
    if (a.final_mode == ZP_FINAL_EPNP_GN) {
        const double fx = cam.fu, fy = cam.fv, cx = cam.uc, cy = cam.vc;
        for (int iter = 0; iter < 5; iter++) {
            double R[9], t[3];
            for (int e = 0; e < 9; e++) R[e] = s_pose[e];
            for (int e = 0; e < 3; e++) t[e] = s_pose[9 + e];
            for (int vl = 0; vl < NR; vl++) {
                for (int q = 0; q < 27; q++) acc[q] = 0;     // 21 JtJ (upper) + 6 Jtr
                const int nv = s_nvr[vl];
                for (int k = tid; k < nv; k += FCL_THREADS) {
                    const int i = inlier_at(vl, k);
                    double X = pX[i], Y = pY[i], Z = pZ[i];
                    double px = R[0] * X + R[1] * Y + R[2] * Z, py = R[3] * X + R[4] * Y + R[5] * Z, pz = R[6] * X + R[7] * Y + R[8] * Z;
                    double xc = px + t[0], yc = py + t[1], zc = pz + t[2], iz = 1.0 / zc;
                    double ru = fx * xc * iz + cx - (double)pu[i], rv = fy * yc * iz + cy - (double)pv[i];
                    double ju[3] = {fx * iz, 0, -fx * xc * iz * iz}, jv[3] = {0, fy * iz, -fy * yc * iz * iz};
                    // cam point = exp(w) (R X) + t + dt  ->  d/dw = -[R X]_x , d/dt = I
                    double Ju[6] = {ju[1] * (-pz) + ju[2] * py, ju[0] * pz + ju[2] * (-px), ju[0] * (-py) + ju[1] * px, ju[0], ju[1], ju[2]};
                    double Jv[6] = {jv[1] * (-pz) + jv[2] * py, jv[0] * pz + jv[2] * (-px), jv[0] * (-py) + jv[1] * px, jv[0], jv[1], jv[2]};
                    int q = 0;
                    for (int r = 0; r < 6; r++)
                        for (int c = r; c < 6; c++) { acc[q] += Ju[r] * Ju[c] + Jv[r] * Jv[c]; q++; }
                    for (int r = 0; r < 6; r++) acc[21 + r] += Ju[r] * ru + Jv[r] * rv;
                }
                block_reduce<27, FCL_THREADS>(acc, s_red, s_part[vl]);
            }
            fcl_total<CL, 27>(s_part, s_sum);
            if (tid == 0) {
                double A[36], g[6], d[6];
                int q = 0;
                for (int r = 0; r < 6; r++)
                    for (int c = r; c < 6; c++) { A[6 * r + c] = s_sum[q]; A[6 * c + r] = s_sum[q]; q++; }
                for (int r = 0; r < 6; r++) g[r] = -s_sum[21 + r];
                bool okc = true;                         // Cholesky solve A d = g
                for (int r = 0; r < 6 && okc; r++) {
                    for (int c = 0; c <= r; c++) {
                        double sacc = A[6 * r + c];
                        for (int k = 0; k < c; k++) sacc -= A[6 * r + k] * A[6 * c + k];
                        if (r == c) { if (sacc <= 0) { okc = false; break; } A[6 * r + r] = sqrt(sacc); }
                        else A[6 * r + c] = sacc / A[6 * c + c];
                    }
                }
                if (okc) {
                    for (int r = 0; r < 6; r++) { double sacc = g[r]; for (int k = 0; k < r; k++) sacc -= A[6 * r + k] * d[k]; d[r] = sacc / A[6 * r + r]; }
                    for (int r = 5; r >= 0; r--) { double sacc = d[r]; for (int k = r + 1; k < 6; k++) sacc -= A[6 * k + r] * d[k]; d[r] = sacc / A[6 * r + r]; }
                    double th = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    double E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                    if (th > 1e-300) {
                        double kx = d[0] / th, ky = d[1] / th, kz = d[2] / th, sn = sin(th), cs = cos(th), oc = 1 - cs;
                        E[0] = cs + kx * kx * oc; E[1] = kx * ky * oc - kz * sn; E[2] = kx * kz * oc + ky * sn;
                        E[3] = ky * kx * oc + kz * sn; E[4] = cs + ky * ky * oc; E[5] = ky * kz * oc - kx * sn;
                        E[6] = kz * kx * oc - ky * sn; E[7] = kz * ky * oc + kx * sn; E[8] = cs + kz * kz * oc;
                    }
                    double Rn[9];
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++) Rn[3 * r + c] = E[3 * r] * R[c] + E[3 * r + 1] * R[3 + c] + E[3 * r + 2] * R[6 + c];
                    for (int e = 0; e < 9; e++) s_pose[e] = Rn[e];
                    for (int e = 0; e < 3; e++) s_pose[9 + e] = t[e] + d[3 + e];
                }
            }
            __syncthreads();
        }
    }
    if (writer) {
        if (tid < 12) out[tid] = s_pose[tid];
        if (a.records && tid < 14) a.records[14 * (size_t)b + tid] = tid < 12 ? s_pose[tid] : tid == 12 ? (double)ni : (double)s_status;
    }
    if (CL > 1) cg::this_cluster().sync();          // no CTA may exit while a peer can still read its shared memory
}

// ---------------------------------------------------------------------------------------------------------------
// FP32 FMA peak probe (roofline denominator for zp_score)
// ---------------------------------------------------------------------------------------------------------------
__global__ void zp_fma_probe_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
int zp_launch_samples(zp_ctx* ctx, const int32_t* counts, int cap, int B, int H, int m, int mode, uint64_t seed,
                      int32_t* samples, cudaStream_t st) {
    ZP_TIME_BEGIN(ctx, st);
    zp_samples_kernel<<<B, SMP_THREADS, 0, st>>>(counts, cap, B, H, m, mode, seed, ctx->d_rng, ctx->n_rng, samples);
    ZP_CHECK_LAUNCH(ctx, "zp_samples_kernel");
    return 0;
}

int zp_launch_minimal_cv(zp_ctx*, const float*, int, const int32_t*, const double*, const int32_t*, int, int, int, int,
                         const int32_t*, const int32_t*, int, float, double*, float*, int32_t*, cudaStream_t);

// hypotheses [h0, h0 + hw) of the crops not flagged in crop_done (nullable).  ctx->solver picks the solver: the exact
// replay of cv2's EPnP (zp_cvsolve.cu, default) or the fast float64 solver of this file.
int zp_launch_minimal(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                      const int32_t* samples, int B, int H, int h0, int hw, const int32_t* crop_done, const int32_t* rs, int m,
                      float thr_px, double* hyp_poses, float* hyp_P, int32_t* hyp_inliers_to_zero, cudaStream_t st) {
    if (ctx->solver == ZP_SOLVER_CV2)
        return zp_launch_minimal_cv(ctx, corr, cap, counts, K, samples, B, H, h0, hw, crop_done, rs, m, thr_px, hyp_poses, hyp_P,
                                    hyp_inliers_to_zero, st);
    const int per_cta = MIN_THREADS / 4;
    int total = B * hw;
    if (!ctx->min_attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_minimal_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIN_SMEM_BYTES));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_minimal_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIN_SMEM_BYTES));
        const char* e = getenv("ZP_MIN_BLOCKS");
        ctx->min_force = e ? atoi(e) : 0;
        ctx->min_attr_set = true;
    }
    const int force = ctx->min_force;
    const int grid = (total + per_cta - 1) / per_cta;
    // register budget (measured, profiles/README.md): a saturating grid runs fastest with the 255-register build (two CTAs
    // per SM, 312 instead of 1220 bytes of spill stores on the FP64 dependency chains: 1267 vs 1396 us at 1024 crops);
    // a grid of a few waves is quantised in favour of three resident CTAs (256 crops: 351 vs 392 us), and at one wave (64 crops)
    // both take the same time alone but the 168-register build leaves
    // room for another lane's kernels on the SM (257 k vs 236 k poses/s with 3 lanes).  ZP_MIN_BLOCKS=2|3 pins it.
    const bool two = force ? force == 2 : grid > 12 * ctx->sm_count;
    ZP_TIME_BEGIN(ctx, st);
    if (two) zp_minimal_kernel<2><<<grid, MIN_THREADS, MIN_SMEM_BYTES, st>>>(corr, cap, counts, K, samples, B, H, h0, hw, crop_done, rs, m, 1.0 / (double)thr_px, hyp_poses, hyp_P, hyp_inliers_to_zero);
    else zp_minimal_kernel<3><<<grid, MIN_THREADS, MIN_SMEM_BYTES, st>>>(corr, cap, counts, K, samples, B, H, h0, hw, crop_done, rs, m, 1.0 / (double)thr_px, hyp_poses, hyp_P, hyp_inliers_to_zero);
    ZP_CHECK_LAUNCH(ctx, "zp_minimal_kernel");
    return 0;
}

int zp_launch_rs_init(zp_ctx* ctx, const int32_t* counts, int cap, int B, int H, int32_t* rs, int32_t* crop_done,
                      int32_t* hyp_inliers_fill, cudaStream_t st) {
    zp_rs_init_kernel<<<(B + 127) / 128, 128, 0, st>>>(counts, cap, B, H, rs, crop_done, hyp_inliers_fill);
    ZP_CHECK_LAUNCH(ctx, "zp_rs_init_kernel");
    return 0;
}

int zp_launch_rs_replay(zp_ctx* ctx, const int32_t* counts, int cap, const int32_t* hyp_inliers, int B, int H, int h0, int h1,
                        int m, double conf, int select_mode, int32_t* rs, int32_t* crop_done, const float* corr, const double* K,
                        const double* hyp_poses, float thr_px, cudaStream_t st) {
    RsExact ex;
    ex.corr = ctx->rs_no_recount == 1 ? nullptr : corr; ex.K = K; ex.hyp_poses = hyp_poses;
    ex.park = ctx->rs_no_recount == 2 ? 0 : 1;
    ex.inv_thr = 1.0f / thr_px; ex.thr2 = (float)((double)thr_px * (double)thr_px);
    ex.tie = crop_done + B;
    ZP_TIME_BEGIN(ctx, st);
    zp_rs_replay_kernel<<<B, RS_THREADS, 0, st>>>(counts, cap, hyp_inliers, B, H, h0, h1, m, conf, select_mode, rs, crop_done, ex);
    ZP_CHECK_LAUNCH(ctx, "zp_rs_replay_kernel");
    return 0;
}

int zp_launch_poses_to_P(zp_ctx* ctx, const double* poses, const double* K, int B, int H, float thr_px, float* hyp_P,
                         cudaStream_t st) {
    zp_poses_to_P_kernel<<<(B * H + 127) / 128, 128, 0, st>>>(poses, K, B, H, 1.0 / (double)thr_px, hyp_P);
    ZP_CHECK_LAUNCH(ctx, "zp_poses_to_P_kernel");
    return 0;
}

// How a tile's hypotheses are cut into queue chunks: about one tile per CTA slot (64 crops: 832 tiles on 740 slots) -> two
// halves, so that the last round is short (measured -12 % against uncut tiles); otherwise one chunk (finer cuts lose more
// to the per-item work than they win in balance: equal thirds / fifths and shrinking plans such as 80 | 40 | 20 | 10 all
// measured within 3 % of the halves at 64 crops, profiles/r2u_score_plans.txt, and 1-6 % behind one chunk at 1024).
static std::vector<int> score_plan(zp_ctx* ctx, int n_items, int slots, int hw) {
    std::vector<int> cls;
    const int req = ctx->score_hchunk;
    int chunk = req > 0 ? req : hw;
    if (req == 0 && n_items <= 2 * slots && hw >= 64) chunk = (hw + 1) / 2;
    if (chunk > SC_HB) chunk = SC_HB;
    if ((hw + chunk - 1) / chunk > SC_MAXCLS) chunk = (hw + SC_MAXCLS - 1) / SC_MAXCLS;
    for (int h = 0; h < hw; h += chunk) cls.push_back(std::min(chunk, hw - h));
    return cls;
}

template <int NG>
static int launch_score_ng(zp_ctx* ctx, ScoreArgs& a, int smem, cudaStream_t st) {
    constexpr int slot = NG == 1 ? 0 : NG == 2 ? 1 : 2;      // per context (= per device), not per process
    if (!ctx->score_per_sm[slot]) {
        int per_sm = 0;
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_score_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ZP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, zp_score_kernel<NG>, SC_GROUP * NG, smem));
        const char* e = getenv("ZP_SCORE_PER_SM");          // tuning aid: resident scoring CTAs per SM (default: what fits)
        if (e && atoi(e) > 0 && atoi(e) < per_sm) per_sm = atoi(e);
        ctx->score_per_sm[slot] = per_sm < 1 ? 1 : per_sm;
    }
    const int per_sm = ctx->score_per_sm[slot];
    int grid = ctx->sm_count * per_sm;
    const std::vector<int> cls = score_plan(ctx, a.n_items, grid, a.hw);
    a.n_cls = (int)cls.size();
    a.cls_off[0] = 0;
    for (int c = 0; c < a.n_cls; c++) a.cls_off[c + 1] = a.cls_off[c] + cls[c];
    for (int c = a.n_cls + 1; c <= SC_MAXCLS; c++) a.cls_off[c] = a.cls_off[a.n_cls];
    if (grid > a.n_items * a.n_cls) grid = a.n_items * a.n_cls;
    ZP_TIME_BEGIN(ctx, st);
    zp_score_kernel<NG><<<grid, SC_GROUP * NG, smem, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_score_kernel");
    return 0;
}

int zp_launch_score(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const float* hyp_P, int B, int H,
                    int h0, int hw, const int32_t* crop_done, const int32_t* rs, float thr_px, int32_t* hyp_inliers,
                    bool already_zeroed, cudaStream_t st) {
    ScoreArgs a;
    a.corr = corr; a.cap = cap; a.counts = counts; a.hyp_P = hyp_P; a.B = B; a.H = H; a.inv_thr = 1.0f / thr_px;
    a.hyp_inliers = hyp_inliers; a.counters = ctx->d_counters; a.h0 = h0; a.hw = hw; a.crop_done = crop_done; a.rs = rs;
    const int max_tiles = (cap + SC_TILE - 1) / SC_TILE;
    a.n_items = B * max_tiles;
    const int smem = 5 * SC_TILE * sizeof(float) + SC_HB * (6 * sizeof(ulonglong2) + sizeof(int));
    if (!already_zeroed) ZP_CUDA(ctx, cudaMemsetAsync(hyp_inliers, 0, (size_t)B * H * sizeof(int32_t), st));
    const int ng = ctx->score_groups ? ctx->score_groups : 1;
    if (ng == 1) return launch_score_ng<1>(ctx, a, smem, st);
    if (ng == 2) return launch_score_ng<2>(ctx, a, smem, st);
    return launch_score_ng<4>(ctx, a, smem, st);
}

int zp_launch_final_split(zp_ctx* ctx, const FinalArgs& a, cudaStream_t st);

int zp_launch_final(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                    const double* hyp_poses, const int32_t* hyp_inliers, int B, int H, int m, double conf, int select_mode,
                    float thr_px, int final_mode, double* poses, int32_t* n_inliers, int32_t* status, int32_t* best_idx,
                    uint8_t* inlier_mask, const int32_t* rs, int32_t* iters_run, double* records, cudaStream_t st) {
    FinalArgs a;
    a.rs = rs; a.iters_run = iters_run; a.records = records;
    a.corr = corr; a.cap = cap; a.counts = counts; a.K = K; a.hyp_poses = hyp_poses; a.hyp_inliers = hyp_inliers;
    a.B = B; a.H = H; a.m = m; a.conf = conf; a.select_mode = select_mode; a.inv_thr = 1.0f / thr_px;
    a.thr2 = (float)((double)thr_px * (double)thr_px);
    a.final_mode = final_mode; a.poses = poses; a.n_inliers = n_inliers; a.status = status; a.best_idx = best_idx;
    a.inlier_mask = inlier_mask;
    if (!ctx->fin_attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_final_cl_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_final_cl_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
        const char* e = getenv("ZP_FINAL_CL");          // tuning aid: force the cluster (4) or the single-CTA (1) form
        if (e && !ctx->fin_form_set) ctx->fin_force = atoi(e);
        ctx->fin_attr_set = true;
    }
    // packed inlier indices: 16 virtual warps x chunk entries per crop (4 per CTA in the cluster form)
    const int chunk = ((cap + FIN_VR * FCL_THREADS - 1) / (FIN_VR * FCL_THREADS)) * 32;
    const size_t idx_bytes = (size_t)FIN_VR * FCL_WARPS * chunk * sizeof(uint16_t);
    if (idx_bytes > 160 * 1024) ZP_FAIL(ctx, -1, "zp_ransac: cap %d needs %zu bytes of shared memory in the final solve", cap, idx_bytes);
    // the cluster form while its 4 B CTAs fit one wave of two CTAs per SM: below that the GPU is not full and a crop's point
    // passes are 4x shorter; above it the single-CTA form does the same arithmetic without the redundant solver phases.
    // Both produce identical bits (same partition, same reduction order).
    // default: the split form (zp_finsplit.cu); the Gauss-Newton polish and explicitly selected forms run here
    if ((ctx->fin_force == 2 || ctx->fin_force == 0) && final_mode == ZP_FINAL_EPNP) return zp_launch_final_split(ctx, a, st);
    const bool cluster = ctx->fin_force == 1 || ctx->fin_force == 4 ? ctx->fin_force == 4 : 4 * B <= 2 * ctx->sm_count;
    ZP_TIME_BEGIN(ctx, st);
    if (cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(4 * B));
        cfg.blockDim = dim3(FCL_THREADS);
        cfg.dynamicSmemBytes = idx_bytes / FIN_VR;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        ZP_CUDA(ctx, cudaLaunchKernelEx(&cfg, zp_final_cl_kernel<4>, a));
    } else {
        zp_final_cl_kernel<1><<<B, FCL_THREADS, idx_bytes, st>>>(a);
    }
    ZP_CHECK_LAUNCH(ctx, "zp_final_kernel");
    return 0;
}

__global__ void zp_fma2_probe_kernel(float* out, int iters, float a, float b) {
    f32x2 pa = zp_pack2(a, a), pb = zp_pack2(b, b);
    f32x2 x0 = zp_pack2(threadIdx.x, 1.f), x1 = zp_pack2(2.f, threadIdx.x), x2 = zp_pack2(3.f, 4.f), x3 = zp_pack2(5.f, 6.f),
          x4 = zp_pack2(7.f, 8.f), x5 = zp_pack2(9.f, 1.5f), x6 = zp_pack2(2.5f, 3.5f), x7 = zp_pack2(4.5f, 5.5f);
    for (int i = 0; i < iters; i++) {
        x0 = zp_fma2(x0, pa, pb); x1 = zp_fma2(x1, pa, pb); x2 = zp_fma2(x2, pa, pb); x3 = zp_fma2(x3, pa, pb);
        x4 = zp_fma2(x4, pa, pb); x5 = zp_fma2(x5, pa, pb); x6 = zp_fma2(x6, pa, pb); x7 = zp_fma2(x7, pa, pb);
    }
    f32x2 s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s ^ (uint32_t)(s >> 32));
}

// FP64 peak probe (roofline denominator of the solver kernels): 8 independent DFMA chains per thread
__global__ void zp_dfma_probe_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int zp_launch_dfma_probe(zp_ctx* ctx, int iters, double* out_tflops) {
    const int blocks = ctx->sm_count * 8, threads = 256;
    double* d = nullptr;
    ZP_CUDA(ctx, cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        zp_dfma_probe_kernel<<<blocks, threads>>>(d, iters, 0.999, 0.001);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        ctx->launches++;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) ZP_FAIL(ctx, -3, "dfma probe failed: %s", cudaGetErrorString(e));
    *out_tflops = best;
    return 0;
}

int zp_read_debug_clocks(long long* host16) {
    return cudaMemcpyFromSymbol(host16, zp_dbg_clk, sizeof(long long) * 24) == cudaSuccess ? 0 : -2;
}

int zp_launch_fma_probe(zp_ctx* ctx, int iters, int packed, double* out_tflops) {
    const int blocks = ctx->sm_count * 8, threads = 256;
    float* d = nullptr;
    ZP_CUDA(ctx, cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        if (packed) zp_fma2_probe_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
        else zp_fma_probe_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        ctx->launches++;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = (packed ? 2.0 : 1.0) * 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) ZP_FAIL(ctx, -3, "fma probe failed: %s", cudaGetErrorString(e));
    *out_tflops = best;
    return 0;
}
