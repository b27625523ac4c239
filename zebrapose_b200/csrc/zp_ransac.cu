// Kernel family 2: batched RANSAC-PnP (replaces cv2.solvePnPRansac(..., reprojectionError=2, iterationsCount=150,
// flags=SOLVEPNP_EPNP) + cv2.Rodrigues, /root/reference/zebrapose/binary_code_helper/CNN_output_to_pose.py:155-158).
//
//   zp_samples_kernel   one CTA per crop, one thread per hypothesis: cv::RNG(0xFFFFFFFFFFFFFFFF) replayed from a
//                       precomputed table of its raw 32-bit outputs; redraws on duplicates shift later hypotheses, which
//                       is resolved by a fixed-point iteration over a block prefix sum (exact; usually 1-2 rounds).
//                       Philox4x32-10 mode is counter based and needs no iteration.
//   zp_minimal_kernel   one thread per hypothesis: float64 EPnP on the m sampled points (12x12 problem interleaved in
//                       shared memory so a warp's accesses are conflict-free, row i of the Jacobi held in registers)
//   zp_score_kernel     FP32-FMA bound: every correspondence x every hypothesis.  Correspondence tiles (SoA planes) are
//                       staged into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered),
//                       hypotheses K[R|t] live in shared memory and are broadcast; the test is division free:
//                       (x - u z)^2 + (y - v z)^2 <= thr^2 z^2.  Counts: per-thread -> warp REDUX -> shared -> global.
//   zp_final_kernel     one CTA per crop: cv2's sequential "strictly greater + RANSACUpdateNumIters" rule replayed over
//                       the H counts (or argmax), then EPnP on all inliers of the winner: block reductions of the 52
//                       EPnP sums, 16-lane cooperative Jacobi for the 12x12 null space, the three beta candidates on
//                       three lanes, optional Gauss-Newton polish of the reprojection error
//
// Algorithmic FP32 work of scoring: 27 flop per (correspondence, hypothesis).
#include "zp_common.cuh"
#include "zp_epnp.cuh"

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

// draws m distinct indices in [0,n) from the raw stream starting at table position `pos`; returns the number of raw
// values consumed, or -1 if the table would be overrun
__device__ __forceinline__ int draw_from_table(const uint32_t* __restrict__ tab, int n_tab, int pos, int n, int m, int* idx) {
    int p = pos;
    for (int j = 0; j < m; j++) {
        int v;
        bool dup;
        do {
            if (p >= n_tab) return -1;
            v = (int)(tab[p++] % (uint32_t)n);
            dup = false;
            for (int q = 0; q < j; q++) dup |= idx[q] == v;
        } while (dup);
        idx[j] = v;
    }
    return p - pos;
}

// grid = B, block = SMP_THREADS; thread h handles hypotheses h, h + SMP_THREADS, ...
__global__ void __launch_bounds__(SMP_THREADS)
zp_samples_kernel(const int32_t* __restrict__ counts, int cap, int B, int H, int m, int mode, uint64_t seed,
                  const uint32_t* __restrict__ rng_tab, int n_tab, int32_t* __restrict__ samples) {
    const int b = blockIdx.x, tid = threadIdx.x;
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
    // cv2 replay.  off[h] = first raw value hypothesis h consumes = m*h + (redraws of all earlier hypotheses).
    __shared__ int s_cons[ZP_MAX_HYPOTHESES];
    __shared__ int s_off[ZP_MAX_HYPOTHESES];
    __shared__ int s_flag;
    for (int h = tid; h < H; h += SMP_THREADS) s_off[h] = m * h;
    __syncthreads();
    bool overflow = false;
    for (int round = 0; round <= H; round++) {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        for (int h = tid; h < H; h += SMP_THREADS) {
            int idx[8];
            int c = draw_from_table(rng_tab, n_tab, s_off[h], n, m, idx);
            if (c < 0) { c = m; atomicOr(&s_flag, 2); }
            s_cons[h] = c;
        }
        __syncthreads();
        if (s_flag & 2) { overflow = true; break; }
        if (tid == 0) {                            // H <= 1024: a serial scan is a few hundred cycles
            int run = 0, chg = 0;
            for (int h = 0; h < H; h++) {
                chg |= s_off[h] != run;
                s_off[h] = run;
                run += s_cons[h];
            }
            if (chg) s_flag = 1;
        }
        __syncthreads();
        if (!(s_flag & 1)) break;
        __syncthreads();
    }
    if (!overflow) {
        for (int h = tid; h < H; h += SMP_THREADS) {
            int idx[8];
            draw_from_table(rng_tab, n_tab, s_off[h], n, m, idx);
            for (int j = 0; j < m; j++) out[h * m + j] = idx[j];
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
constexpr int MIN_THREADS = 64;
constexpr int ZP_MAX_M = 8;

__global__ void __launch_bounds__(MIN_THREADS)
zp_minimal_kernel(const float* __restrict__ corr, int cap, const int32_t* __restrict__ counts,
                  const double* __restrict__ Kmat, const int32_t* __restrict__ samples, int B, int H, int m,
                  double* __restrict__ hyp_poses) {
    extern __shared__ double smem_d[];
    const int g = blockIdx.x * MIN_THREADS + threadIdx.x;
    if (g >= B * H) return;
    const int b = g / H;
    double* out = hyp_poses + (size_t)g * 12;
    const int32_t* sidx = samples + (size_t)g * m;
    const int n = min(counts[b], cap);
    bool valid = n >= m;
    for (int j = 0; j < m; j++) valid = valid && sidx[j] >= 0 && sidx[j] < n;
    if (!valid) {
        for (int e = 0; e < 12; e++) out[e] = nan("");
        return;
    }
    const double* Kb = Kmat + 9 * (size_t)b;
    const ZpCam cam{Kb[0], Kb[4], Kb[2], Kb[5]};
    const float* cb = corr + (size_t)b * 5 * cap;
    double X[ZP_MAX_M], Y[ZP_MAX_M], Z[ZP_MAX_M], xn[ZP_MAX_M], yn[ZP_MAX_M];   // xn, yn: pixel coordinates
    double c0[3] = {0, 0, 0};
    for (int j = 0; j < m; j++) {
        int i = sidx[j];
        xn[j] = (double)cb[i];
        yn[j] = (double)cb[cap + i];
        X[j] = cb[2 * (size_t)cap + i]; Y[j] = cb[3 * (size_t)cap + i]; Z[j] = cb[4 * (size_t)cap + i];
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
    ZpSums sums;
    for (int q = 0; q < 10; q++) { sums.s0[q] = 0; sums.sx[q] = 0; sums.sy[q] = 0; sums.sr[q] = 0; }
    for (int q = 0; q < 12; q++) sums.w[q] = 0;
    sums.n = m;
    double a_first[4];
    for (int j = 0; j < m; j++) {
        double a[4];
        zp_alphas(cp, X[j], Y[j], Z[j], a);
        if (j == 0) { a_first[0] = a[0]; a_first[1] = a[1]; a_first[2] = a[2]; a_first[3] = a[3]; }
        zp_accumulate(sums, a, cam.uc - xn[j], cam.vc - yn[j], X[j] - c0[0], Y[j] - c0[1], Z[j] - c0[2]);
    }
    ZpMat At{smem_d + threadIdx.x, MIN_THREADS};
    zp_nullspace_serial(At, sums, cam);            // rows 0..3 of At now hold the null-space vectors
    double L[60], rho[6];
    zp_L_rho(At, cp, L, rho);
    bool have = false;
    double best_err = 0, Rb[9], tb[3];
    for (int c = 0; c < 3; c++) {
        double R[9], t[3];
        if (!zp_candidate(c, L, rho, At, sums, a_first, c0, R, t)) continue;
        double e = 0;
        for (int j = 0; j < m; j++) e += zp_reproj_dist(R, t, cam, X[j], Y[j], Z[j], xn[j], yn[j]);
        e /= m;
        if (!(e == e)) continue;
        if (!have || e < best_err) {
            have = true; best_err = e;
            for (int q = 0; q < 9; q++) Rb[q] = R[q];
            for (int q = 0; q < 3; q++) tb[q] = t[q];
        }
    }
    if (!have) {
        for (int e = 0; e < 12; e++) out[e] = nan("");
        return;
    }
    for (int e = 0; e < 9; e++) out[e] = Rb[e];
    for (int e = 0; e < 3; e++) out[9 + e] = tb[e];
}

// ---------------------------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_PPT = 4;                          // correspondences per thread
constexpr int SC_TILE = SC_THREADS * SC_PPT;       // correspondences per tile
constexpr int SC_STAGES = 2;

// projection rows in float32 from a float64 pose: P = K [R|t] evaluated in double, rounded once
__device__ __forceinline__ void zp_make_P(const double* pose, const double* K, float P[12]) {
    const double fx = K[0], sk = K[1], cx = K[2], fy = K[4], cy = K[5];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        double r0 = c < 3 ? pose[c] : pose[9], r1 = c < 3 ? pose[3 + c] : pose[10], r2 = c < 3 ? pose[6 + c] : pose[11];
        P[c] = (float)(fx * r0 + sk * r1 + cx * r2);
        P[4 + c] = (float)(fy * r1 + cy * r2);
        P[8 + c] = (float)r2;
    }
}

// the inlier test, shared verbatim by scoring and the final solve (explicit fmaf so both kernels round identically)
__device__ __forceinline__ bool zp_is_inlier(const float4& p0, const float4& p1, const float4& p2, float u, float v,
                                             float X, float Y, float Z, float thr2) {
    float x = fmaf(p0.x, X, fmaf(p0.y, Y, fmaf(p0.z, Z, p0.w)));
    float y = fmaf(p1.x, X, fmaf(p1.y, Y, fmaf(p1.z, Z, p1.w)));
    float z = fmaf(p2.x, X, fmaf(p2.y, Y, fmaf(p2.z, Z, p2.w)));
    float dx = fmaf(-u, z, x);
    float dy = fmaf(-v, z, y);
    float e = fmaf(dy, dy, __fmul_rn(dx, dx));
    float lim = __fmul_rn(__fmul_rn(z, z), thr2);
    return e <= lim;                               // NaN poses compare false -> 0 inliers
}

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

struct ScoreArgs {
    const float* corr; int cap; const int32_t* counts; const double* K; const double* hyp_poses;
    int B, H; float thr2; int32_t* hyp_inliers; int nsplit;
};

// grid = (nsplit, B).  CTA (split s of crop b) walks tiles s, s+nsplit, ... of the crop's correspondences.
__global__ void __launch_bounds__(SC_THREADS) zp_score_kernel(ScoreArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_pts = (float*)smem_raw;                                   // [SC_STAGES][5][SC_TILE]
    float4* s_P = (float4*)(s_pts + SC_STAGES * 5 * SC_TILE);          // [H][3]
    int* s_cnt = (int*)(s_P + 3 * a.H);                                // [H]
    __shared__ __align__(8) uint64_t s_bar[SC_STAGES];

    const int b = blockIdx.y, split = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = min(a.counts[b], a.cap);
    const int n_tiles = (n + SC_TILE - 1) / SC_TILE;
    const int H = a.H;
    const bool direct = a.nsplit == 1;
    if (split >= n_tiles && !(direct || split == 0)) return;           // nothing to do (outputs pre-zeroed)

    if (tid == 0) {
        for (int s = 0; s < SC_STAGES; s++) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int h = tid; h < H; h += SC_THREADS) {
        float P[12];
        zp_make_P(a.hyp_poses + ((size_t)b * H + h) * 12, a.K + 9 * (size_t)b, P);
        s_P[3 * h + 0] = make_float4(P[0], P[1], P[2], P[3]);
        s_P[3 * h + 1] = make_float4(P[4], P[5], P[6], P[7]);
        s_P[3 * h + 2] = make_float4(P[8], P[9], P[10], P[11]);
        s_cnt[h] = 0;
    }
    __syncthreads();

    const float* cb = a.corr + (size_t)b * 5 * a.cap;
    auto issue = [&](int tile, int stage) {
        int start = tile * SC_TILE;
        int cnt = min(SC_TILE, n - start);
        uint32_t bytes = (uint32_t)((cnt + 3) & ~3) * 4u;              // 16-byte granules; cap % 4 == 0 keeps it in bounds
        mbar_expect_tx(&s_bar[stage], 5 * bytes);
        for (int pl = 0; pl < 5; pl++)
            tma_load_1d(s_pts + (stage * 5 + pl) * SC_TILE, cb + (size_t)pl * a.cap + start, bytes, &s_bar[stage]);
    };
    if (tid == 0 && split < n_tiles) issue(split, 0);

    int it = 0;
    for (int tile = split; tile < n_tiles; tile += a.nsplit, it++) {
        const int stage = it & 1;
        if (tid == 0 && tile + a.nsplit < n_tiles) issue(tile + a.nsplit, stage ^ 1);   // prefetch next tile
        mbar_wait(&s_bar[stage], (it >> 1) & 1);
        const float* tp = s_pts + stage * 5 * SC_TILE;
        const int start = tile * SC_TILE;
        float u[SC_PPT], v[SC_PPT], X[SC_PPT], Y[SC_PPT], Z[SC_PPT];
        bool live[SC_PPT];
#pragma unroll
        for (int j = 0; j < SC_PPT; j++) {
            int i = tid + j * SC_THREADS;
            live[j] = start + i < n;
            u[j] = tp[i]; v[j] = tp[SC_TILE + i]; X[j] = tp[2 * SC_TILE + i]; Y[j] = tp[3 * SC_TILE + i];
            Z[j] = tp[4 * SC_TILE + i];
            if (!live[j]) { u[j] = 0.f; v[j] = 0.f; X[j] = 0.f; Y[j] = 0.f; Z[j] = __int_as_float(0x7fc00000); }  // NaN -> never an inlier
        }
#pragma unroll 2
        for (int h = 0; h < H; h++) {
            const float4 p0 = s_P[3 * h], p1 = s_P[3 * h + 1], p2 = s_P[3 * h + 2];
            int c = 0;
#pragma unroll
            for (int j = 0; j < SC_PPT; j++) c += zp_is_inlier(p0, p1, p2, u[j], v[j], X[j], Y[j], Z[j], a.thr2) ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0 && c) atomicAdd(&s_cnt[h], c);
        }
        __syncthreads();        // everyone is done with this stage before it is refilled two iterations later
    }
    __syncthreads();
    int32_t* out = a.hyp_inliers + (size_t)b * H;
    for (int h = tid; h < H; h += SC_THREADS) {
        if (direct) out[h] = s_cnt[h];
        else if (s_cnt[h]) atomicAdd(&out[h], s_cnt[h]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// winner selection + final solve on the inliers of the winner: one CTA per crop
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

constexpr int FIN_THREADS = 256;

// phase timestamps of CTA 0 of the last zp_final_kernel launch (debug / profiling aid, read with zp_debug_clocks)
__device__ long long zp_dbg_clk[16];
#define ZP_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) zp_dbg_clk[i] = clock64(); } while (0)

template <int NV>
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

struct FinalArgs {
    const float* corr; int cap; const int32_t* counts; const double* K; const double* hyp_poses;
    const int32_t* hyp_inliers; int B, H, m; double conf; int select_mode; float thr2; int final_mode;
    double* poses; int32_t* n_inliers; int32_t* status; int32_t* best_idx; uint8_t* inlier_mask;
};

__global__ void __launch_bounds__(FIN_THREADS) zp_final_kernel(FinalArgs a) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    __shared__ double s_red[(FIN_THREADS / 32) * 52];
    __shared__ double s_sum[52];
    __shared__ double s_V[48];
    __shared__ ZpControl s_cp;
    __shared__ ZpSums s_sums;
    __shared__ double s_candR[3][9], s_candt[3][3];
    __shared__ int s_candok[3];
    __shared__ double s_pose[12];
    __shared__ int s_first, s_n, s_best, s_status;
    extern __shared__ uint32_t s_dyn[];               // inlier bitset (cap+31)/32 words, then H counts
    uint32_t* s_mask = s_dyn;
    int* s_hi = (int*)(s_dyn + (a.cap + 31) / 32);

    ZP_STAMP(0);
    double* out = a.poses + 12 * (size_t)b;
    const int n_raw = a.counts[b];
    const int n = min(n_raw, a.cap);
    // ---- winner: cv2's rule replayed over the H counts (PnPRansac / RANSACPointSetRegistrator::run)
    for (int h = tid; h < a.H; h += FIN_THREADS) s_hi[h] = a.hyp_inliers[(size_t)b * a.H + h];
    __syncthreads();
    if (tid == 0) {
        int best = -1, st = ZP_OK;
        if (n_raw == 0) st = ZP_NO_MASK_PIXELS;
        else if (n < 6) st = ZP_TOO_FEW_POINTS;       // CNN_output_to_pose.py:126
        else {
            int maxgood = 0;
            if (a.select_mode == ZP_SELECT_CV2_REPLAY) {
                int niters = max(a.H, 1);
                for (int it = 0; it < niters && it < a.H; it++) {
                    int good = s_hi[it];
                    if (good > max(maxgood, a.m - 1)) {
                        best = it; maxgood = good;
                        niters = zp_update_iters(a.conf, (double)(n - good) / n, a.m, niters);
                    }
                }
            } else {
                for (int it = 0; it < a.H; it++)
                    if (s_hi[it] > max(maxgood, a.m - 1)) { best = it; maxgood = s_hi[it]; }
            }
            if (best < 0) st = ZP_RANSAC_NO_MODEL;
        }
        s_best = best; s_status = st;
        a.status[b] = st;
        if (a.best_idx) a.best_idx[b] = best;
        s_first = 0x7fffffff; s_n = 0;
    }
    for (int i = tid; i < (a.cap + 31) / 32; i += FIN_THREADS) s_mask[i] = 0;
    __syncthreads();
    const int best = s_best;
    if (best < 0) {     // no model: cv2 leaves rvec = tvec = 0 and the reference reports R = I, t = 0
        if (tid < 12) out[tid] = (tid == 0 || tid == 4 || tid == 8) ? 1.0 : 0.0;
        if (tid == 0) a.n_inliers[b] = 0;
        if (a.inlier_mask)
            for (int i = tid; i < a.cap; i += FIN_THREADS) a.inlier_mask[(size_t)b * a.cap + i] = 0;
        return;
    }
    const float* cb = a.corr + (size_t)b * 5 * a.cap;
    const float *pu = cb, *pv = cb + a.cap, *pX = cb + 2 * (size_t)a.cap, *pY = cb + 3 * (size_t)a.cap, *pZ = cb + 4 * (size_t)a.cap;
    const double* Kb = a.K + 9 * (size_t)b;
    const double* hp = a.hyp_poses + ((size_t)b * a.H + best) * 12;
    float P[12];
    zp_make_P(hp, Kb, P);
    const float4 p0 = make_float4(P[0], P[1], P[2], P[3]), p1 = make_float4(P[4], P[5], P[6], P[7]),
                 p2 = make_float4(P[8], P[9], P[10], P[11]);
    ZP_STAMP(1);
    // ---- pass 0: inlier set of the winner (same predicate as zp_score_kernel), centroid
    double acc[52];
    for (int q = 0; q < 52; q++) acc[q] = 0;
    int my_n = 0, my_first = 0x7fffffff;
    for (int i0 = 0; i0 < n; i0 += FIN_THREADS) {     // warp-aligned so the bitset is built with ballots
        int i = i0 + tid;
        bool in = i < n && zp_is_inlier(p0, p1, p2, pu[i], pv[i], pX[i], pY[i], pZ[i], a.thr2);
        unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0 && i < a.cap) s_mask[i >> 5] = bal;
        if (a.inlier_mask && i < n) a.inlier_mask[(size_t)b * a.cap + i] = in;
        if (in) {
            my_n++; my_first = min(my_first, i);
            acc[0] += pX[i]; acc[1] += pY[i]; acc[2] += pZ[i];
        }
    }
    if (a.inlier_mask)
        for (int i = n + tid; i < a.cap; i += FIN_THREADS) a.inlier_mask[(size_t)b * a.cap + i] = 0;
    my_n = __reduce_add_sync(0xffffffffu, my_n);
    my_first = __reduce_min_sync(0xffffffffu, my_first);
    if (lane == 0) { atomicAdd(&s_n, my_n); atomicMin(&s_first, my_first); }
    block_reduce<3>(acc, s_red, s_sum);
    const int ni = s_n;
    if (tid == 0) a.n_inliers[b] = ni;
    if (ni < 4) {       // cannot happen after selection (good > m-1 >= 3) but keep the output defined
        if (tid < 12) out[tid] = hp[tid];
        return;
    }
    const double c0[3] = {s_sum[0] / ni, s_sum[1] / ni, s_sum[2] / ni};
    __syncthreads();
    ZP_STAMP(2);
    // ---- pass 1: scatter matrix
    for (int q = 0; q < 9; q++) acc[q] = 0;
    for (int i = tid; i < n; i += FIN_THREADS)
        if (s_mask[i >> 5] >> (i & 31) & 1u) {
            double d0 = pX[i] - c0[0], d1 = pY[i] - c0[1], d2 = pZ[i] - c0[2];
            acc[0] = fma(d0, d0, acc[0]); acc[1] = fma(d0, d1, acc[1]); acc[2] = fma(d0, d2, acc[2]);
            acc[4] = fma(d1, d1, acc[4]); acc[5] = fma(d1, d2, acc[5]); acc[8] = fma(d2, d2, acc[8]);
        }
    acc[3] = acc[1]; acc[6] = acc[2]; acc[7] = acc[5];
    block_reduce<9>(acc, s_red, s_sum);
    ZP_STAMP(3);
    if (tid == 0) {
        double C[9];
        for (int q = 0; q < 9; q++) C[q] = s_sum[q];
        zp_control_points(c0, C, (double)ni, s_cp);
    }
    __syncthreads();
    ZP_STAMP(4);
    // ---- pass 2: the 52 EPnP sums
    const ZpCam cam{Kb[0], Kb[4], Kb[2], Kb[5]};
    {
        ZpSums s;
        for (int q = 0; q < 10; q++) { s.s0[q] = 0; s.sx[q] = 0; s.sy[q] = 0; s.sr[q] = 0; }
        for (int q = 0; q < 12; q++) s.w[q] = 0;
        const ZpControl cp = s_cp;
        for (int i = tid; i < n; i += FIN_THREADS)
            if (s_mask[i >> 5] >> (i & 31) & 1u) {
                double X = pX[i], Y = pY[i], Z = pZ[i];
                double al[4];
                zp_alphas(cp, X, Y, Z, al);
                zp_accumulate(s, al, cam.uc - (double)pu[i], cam.vc - (double)pv[i], X - c0[0], Y - c0[1], Z - c0[2]);
            }
        for (int q = 0; q < 10; q++) { acc[q] = s.s0[q]; acc[10 + q] = s.sx[q]; acc[20 + q] = s.sy[q]; acc[30 + q] = s.sr[q]; }
        for (int q = 0; q < 12; q++) acc[40 + q] = s.w[q];
    }
    block_reduce<52>(acc, s_red, s_sum);
    if (tid < 10) { s_sums.s0[tid] = s_sum[tid]; s_sums.sx[tid] = s_sum[10 + tid]; s_sums.sy[tid] = s_sum[20 + tid]; s_sums.sr[tid] = s_sum[30 + tid]; }
    if (tid < 12) s_sums.w[tid] = s_sum[40 + tid];
    if (tid == 0) s_sums.n = ni;
    __syncthreads();
    ZP_STAMP(5);
    // ---- 12x12 null space on warp 0 (16-lane cooperative Jacobi; both half-warps run the same problem)
    if (tid < 32) {
        const int g = lane & 15;
        double col[12], W[12];
#pragma unroll
        for (int r = 0; r < 12; r++) col[r] = g < 12 ? zp_mtm(s_sums, cam, r, g) : 0.0;
        zp_jacobi12_coop(col, W, g);
        ZP_STAMP(6);
        bool used[12];
#pragma unroll
        for (int r = 0; r < 12; r++) used[r] = false;
        for (int q = 0; q < 4; q++) {
            int bi = -1;
            double bw = 0, bv = 0;
#pragma unroll
            for (int r = 11; r >= 0; r--)
                if (!used[r] && (bi < 0 || W[r] < bw)) { bi = r; bw = W[r]; bv = col[r]; }
#pragma unroll
            for (int r = 0; r < 12; r++) used[r] = used[r] || r == bi;
            if (lane < 12) s_V[q * 12 + g] = bw > ZP_DBL_MIN ? bv / bw : 0.0;
        }
        __syncwarp();
        ZP_STAMP(7);
        // ---- the three beta candidates on three lanes
        if (lane < 3) {
            ZpMat V{s_V, 1};
            double L[60], rho[6], af[4];
            zp_L_rho(V, s_cp, L, rho);
            int f = s_first;
            zp_alphas(s_cp, pX[f], pY[f], pZ[f], af);
            s_candok[lane] = zp_candidate(lane, L, rho, V, s_sums, af, c0, s_candR[lane], s_candt[lane]) ? 1 : 0;
        }
    }
    __syncthreads();
    ZP_STAMP(8);
    // ---- pass 3: mean reprojection distance of the three candidates, pick the best
    for (int q = 0; q < 3; q++) acc[q] = 0;
    for (int i = tid; i < n; i += FIN_THREADS)
        if (s_mask[i >> 5] >> (i & 31) & 1u) {
            double X = pX[i], Y = pY[i], Z = pZ[i], u = pu[i], v = pv[i];
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (s_candok[c]) acc[c] += zp_reproj_dist(s_candR[c], s_candt[c], cam, X, Y, Z, u, v);
        }
    block_reduce<3>(acc, s_red, s_sum);
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
            for (int q = 0; q < 27; q++) acc[q] = 0;     // 21 JtJ (upper) + 6 Jtr
            for (int i = tid; i < n; i += FIN_THREADS)
                if (s_mask[i >> 5] >> (i & 31) & 1u) {
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
            block_reduce<27>(acc, s_red, s_sum);
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
    ZP_STAMP(9);
    if (tid < 12) out[tid] = s_pose[tid];
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
    zp_samples_kernel<<<B, SMP_THREADS, 0, st>>>(counts, cap, B, H, m, mode, seed, ctx->d_rng, ctx->n_rng, samples);
    ZP_CHECK_LAUNCH(ctx, "zp_samples_kernel");
    return 0;
}

int zp_launch_minimal(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                      const int32_t* samples, int B, int H, int m, double* hyp_poses, cudaStream_t st) {
    static bool attr_set = false;
    const int smem = MIN_THREADS * 144 * sizeof(double);
    if (!attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_minimal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    int total = B * H;
    zp_minimal_kernel<<<(total + MIN_THREADS - 1) / MIN_THREADS, MIN_THREADS, smem, st>>>(corr, cap, counts, K, samples, B, H, m, hyp_poses);
    ZP_CHECK_LAUNCH(ctx, "zp_minimal_kernel");
    return 0;
}

int zp_launch_score(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                    const double* hyp_poses, int B, int H, float thr_px, int32_t* hyp_inliers, cudaStream_t st) {
    static int attr_smem = 0;
    ScoreArgs a;
    a.corr = corr; a.cap = cap; a.counts = counts; a.K = K; a.hyp_poses = hyp_poses; a.B = B; a.H = H;
    a.thr2 = thr_px * thr_px; a.hyp_inliers = hyp_inliers;
    const int max_tiles = (cap + SC_TILE - 1) / SC_TILE;
    // enough CTAs to fill the chip a few times over; one CTA per crop once the batch alone does that
    int nsplit = 1;
    while (nsplit < max_tiles && (long)B * nsplit < 4L * ctx->sm_count) nsplit <<= 1;
    if (nsplit > max_tiles) nsplit = max_tiles;
    a.nsplit = nsplit;
    const int smem = SC_STAGES * 5 * SC_TILE * sizeof(float) + H * (3 * sizeof(float4) + sizeof(int));
    if (smem > attr_smem) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    if (nsplit > 1) ZP_CUDA(ctx, cudaMemsetAsync(hyp_inliers, 0, (size_t)B * H * sizeof(int32_t), st));
    zp_score_kernel<<<dim3(nsplit, B), SC_THREADS, smem, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_score_kernel");
    return 0;
}

int zp_launch_final(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                    const double* hyp_poses, const int32_t* hyp_inliers, int B, int H, int m, double conf, int select_mode,
                    float thr_px, int final_mode, double* poses, int32_t* n_inliers, int32_t* status, int32_t* best_idx,
                    uint8_t* inlier_mask, cudaStream_t st) {
    FinalArgs a;
    a.corr = corr; a.cap = cap; a.counts = counts; a.K = K; a.hyp_poses = hyp_poses; a.hyp_inliers = hyp_inliers;
    a.B = B; a.H = H; a.m = m; a.conf = conf; a.select_mode = select_mode; a.thr2 = thr_px * thr_px;
    a.final_mode = final_mode; a.poses = poses; a.n_inliers = n_inliers; a.status = status; a.best_idx = best_idx;
    a.inlier_mask = inlier_mask;
    size_t smem = ((size_t)(cap + 31) / 32 + H) * sizeof(uint32_t);
    zp_final_kernel<<<B, FIN_THREADS, smem, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_final_kernel");
    return 0;
}

int zp_read_debug_clocks(long long* host16) {
    return cudaMemcpyFromSymbol(host16, zp_dbg_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : -2;
}

int zp_launch_fma_probe(zp_ctx* ctx, int iters, double* out_tflops) {
    const int blocks = ctx->sm_count * 8, threads = 256;
    float* d = nullptr;
    ZP_CUDA(ctx, cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        zp_fma_probe_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
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
    if (e != cudaSuccess) ZP_FAIL(ctx, -3, "fma probe failed: %s", cudaGetErrorString(e));
    *out_tflops = best;
    return 0;
}
