// Final solve on the winner's inliers, SPLIT form (default): three kernels instead of one CTA (or cluster) per crop.
// Replaces the last step of cv2.solvePnPRansac -- EPnP on all inliers of the best hypothesis
// (/root/reference/zebrapose/binary_code_helper/CNN_output_to_pose.py:155-157) -- like zp_final_cl_kernel (zp_ransac.cu).
//
// Why split: in the one-kernel forms the point passes (inlier set, scatter matrix, 52 EPnP sums, candidate errors: 83 of
// 127 us per crop, profiles/r2k_final_phases.txt) share a CTA -- and its 255-register allocation -- with the serial solver
// chain (PCA, 12x12 null space, three beta candidates: 43 us on one warp), so a crop's 12 k points were walked by 4 warps
// at 8 warps per SM: FP64 latency, 2-8 % of the DFMA rate.  Here
//   zp_fin_moments_kernel   4 CTAs x 8 warps per crop, <= 128 registers: inlier set of the winner (same predicate as
//                           scoring, doubtful points by cv2's arithmetic), packed inlier lists, and ONE pass of raw
//                           moments T_f = sum f [X Y Z 1][X Y Z 1]^T, f in {1, x, y, x^2+y^2} (x = uc - u, y = vc - v): 40
//                           sums that do not depend on the control points
//   zp_fin_solve_kernel     a warp per crop: centroid, scatter, PCA, control points from T_1; the barycentric coordinates
//                           are affine in the point, alpha = A [X Y Z 1]^T, so EPnP's 52 sums are the contractions
//                           A T_f A^T (and A T_1 for the alignment sums) -- no second and third pass over the points;
//                           then the 12x12 null space on 16 lanes and the three beta candidates on three lanes
//   zp_fin_errors_kernel    4 CTAs x 8 warps per crop: mean reprojection distance of the three candidates over the packed
//                           lists; the last CTA of a crop to finish picks the candidate and writes the pose
// Points are taken relative to the crop's first 3D point (a point of the object), which keeps the cancellation in
// T_1 - n c c^T at the object's own scale wherever the model's origin lies (measured against direct sums: 1e-14 relative
// for centred models).  Partition (32 lists per crop, sized by the crop's own count) and reduction orders are fixed, so a
// crop's pose does not depend on the batch it came in.  The Gauss-Newton polish (final = "epnp+gn") and the forms
// zp_set_final_form selects explicitly stay in zp_final_cl_kernel.
#include <climits>
#include "zp_common.cuh"
#include "zp_epnp.cuh"
#include "zp_proj.cuh"

constexpr int FS_SEG = 4;                       // CTAs per crop in the two point kernels
constexpr int FS_THREADS = 256;
constexpr int FS_WARPS = FS_THREADS / 32;
constexpr int FS_LISTS = FS_SEG * FS_WARPS;     // packed inlier lists per crop (one per warp)
constexpr int FS_NQ = 44;                       // per CTA: 40 moments | inlier count | first inlier index | pad
constexpr int FS_CAND = 40;                     // per crop: 3 x (R[9] t[3] ok) + pad

struct FsWs {
    uint16_t* idx;        // [B][cap]: list l of crop b at idx[b * cap + l * chunk], chunk = the list's own point range
    int32_t* wcnt;        // [B][FS_LISTS]
    double* part;         // [B][FS_SEG][FS_NQ]
    double* cand;         // [B][FS_CAND]
    double* err;          // [B][FS_SEG][4]
    int32_t* done;        // [B] CTAs of the crop that have finished the error pass
    int32_t* state;       // [B] 0: solve, 1: outputs already written (no model / too few points)
    unsigned long long* dbg;   // profiling aid (zp_debug_buffer): clock64 stamps of crop 0 in zp_fin_solve_kernel, slots 8..13
};

__device__ __forceinline__ int fs_chunk(int n) { return ((n + FS_LISTS * 32 - 1) / (FS_LISTS * 32)) * 32; }

__global__ void __launch_bounds__(FS_THREADS, 2) zp_fin_moments_kernel(FinalArgs a, FsWs w) {
    const int b = blockIdx.x / FS_SEG, rank = blockIdx.x % FS_SEG;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    extern __shared__ __align__(8) unsigned char s_dyn[];
    uint16_t* s_idx = (uint16_t*)s_dyn;                  // [FS_WARPS][chunk]
    __shared__ double s_red[FS_WARPS][FS_NQ];
    const int n_raw = a.counts[b];
    const int n = min(n_raw, a.cap);
    int best = -1, st = ZP_OK;
    if (n_raw == 0) st = ZP_NO_MASK_PIXELS;
    else if (n < 6) st = ZP_TOO_FEW_POINTS;               // CNN_output_to_pose.py:126
    else {
        best = a.rs[4 * b + 2];
        if (best < 0) st = ZP_RANSAC_NO_MODEL;
    }
    if (rank == 0 && tid == 0) {
        a.status[b] = st;
        if (a.best_idx) a.best_idx[b] = best;
        if (a.iters_run) a.iters_run[b] = n < 6 ? 0 : a.rs[4 * b + 3];
        w.state[b] = best < 0 ? 1 : 0;
        w.done[b] = 0;
    }
    if (best < 0) {     // no model: cv2 leaves rvec = tvec = 0 and the reference reports R = I, t = 0
        if (rank == 0) {
            if (tid < 12) a.poses[12 * (size_t)b + tid] = (tid == 0 || tid == 4 || tid == 8) ? 1.0 : 0.0;
            if (tid == 0) a.n_inliers[b] = 0;
            if (a.records && tid < 14)
                a.records[14 * (size_t)b + tid] = tid < 12 ? ((tid == 0 || tid == 4 || tid == 8) ? 1.0 : 0.0) : tid == 12 ? 0.0 : (double)st;
        }
        if (a.inlier_mask)
            for (int i = rank * FS_THREADS + tid; i < a.cap; i += FS_SEG * FS_THREADS) a.inlier_mask[(size_t)b * a.cap + i] = 0;
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
    const int chunk = fs_chunk(n);
    const int list = rank * FS_WARPS + warp;
    const int base = list * chunk;                        // this warp's points: [base, base + chunk)
    uint16_t* seg = s_idx + (size_t)warp * chunk;
    // ---- the inlier set of the winner, packed per warp
    int wcount = 0, my_first = INT_MAX;
    for (int j0 = 0; j0 < chunk; j0 += 128) {             // four 32-point groups per trip, their 20 loads up front
        float fu[4], fv[4], fX[4], fY[4], fZ[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = base + j0 + 32 * q + lane;
            const bool ld = j0 + 32 * q < chunk && i < n;
            fu[q] = ld ? pu[i] : 0.f; fv[q] = ld ? pv[i] : 0.f; fX[q] = ld ? pX[i] : 0.f; fY[q] = ld ? pY[i] : 0.f; fZ[q] = ld ? pZ[i] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (j0 + 32 * q >= chunk) break;              // warp-uniform
            const int i = base + j0 + 32 * q + lane;
            bool in = false;
            if (i < n) {
                const float d = zp_inlier_d(p0, p1, p2, fu[q] * a.inv_thr, fv[q] * a.inv_thr, fX[q], fY[q], fZ[q]);
                in = __float_as_int(d) < 0;
                const float z = fmaf(p2.x, fX[q], fmaf(p2.y, fY[q], fmaf(p2.z, fZ[q], p2.w)));
                if (fabsf(d) <= 1e-3f * z * z)            // within ~1e-3 px of the threshold: cv2's own arithmetic decides
                    in = zp_inlier_exact(hp, Kb[0], Kb[4], Kb[2], Kb[5], fu[q], fv[q], fX[q], fY[q], fZ[q], thr2);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (a.inlier_mask && i < n) a.inlier_mask[(size_t)b * a.cap + i] = in;
            if (in) {
                seg[wcount + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
                my_first = min(my_first, i);
            }
            wcount += __popc(bal);
        }
    }
    if (a.inlier_mask)
        for (int i = n + rank * FS_THREADS + tid; i < a.cap; i += FS_SEG * FS_THREADS) a.inlier_mask[(size_t)b * a.cap + i] = 0;
    my_first = __reduce_min_sync(0xffffffffu, my_first);
    __syncwarp();
    {
        uint16_t* gl = w.idx + (size_t)b * a.cap + base;
        for (int k = lane; k < wcount; k += 32) gl[k] = seg[k];
        if (lane == 0) w.wcnt[b * FS_LISTS + list] = wcount;
    }
    // ---- raw moments over the packed list (every lane busy): T[f][q], q = packed (a <= b) of [X Y Z 1]
    double T1[9], Tx[10], Ty[10], Tr[10];
#pragma unroll
    for (int q = 0; q < 10; q++) { if (q < 9) T1[q] = 0; Tx[q] = 0; Ty[q] = 0; Tr[q] = 0; }
    const double g0 = pX[0], g1 = pY[0], g2 = pZ[0];       // pivot: the crop's first 3D point
    const double uc = Kb[2], vc = Kb[5];
    {
        int k = lane;
        bool in = k < wcount;
        int i = in ? seg[k] : 0;
        float fX = in ? pX[i] : 0.f, fY = in ? pY[i] : 0.f, fZ = in ? pZ[i] : 0.f, fu = in ? pu[i] : 0.f, fv = in ? pv[i] : 0.f;
        while (in) {
            const int k2 = k + 32;
            const bool in2 = k2 < wcount;
            const int i2 = in2 ? seg[k2] : 0;
            const float gX = in2 ? pX[i2] : 0.f, gY = in2 ? pY[i2] : 0.f, gZ = in2 ? pZ[i2] : 0.f, gu = in2 ? pu[i2] : 0.f, gv = in2 ? pv[i2] : 0.f;
            {
                zp_moment_add(T1, Tx, Ty, Tr, (double)fX - g0, (double)fY - g1, (double)fZ - g2, uc - (double)fu, vc - (double)fv);
            }
            k = k2; in = in2; fX = gX; fY = gY; fZ = gZ; fu = gu; fv = gv;
        }
    }
    // ---- warp totals -> CTA totals.  The xor butterfly over the lanes (16, 8, 4, 2, 1) in its transposed form: at every level a
    // lane keeps the half of its values that its lane bit selects and adds the partner's copy of that half, so the 40 sums
    // (padded to 64) cost 32 + 16 + 8 + 4 + 2 exchanges instead of 5 x 40 -- the plain butterfly spent as many FP64 adds
    // here as the moments themselves.  Same pairs at every level, hence the same bits.  Lane l ends with values 2l, 2l + 1.
    {
        double v[64];
#pragma unroll
        for (int q = 0; q < 64; q++) v[q] = q < 9 ? T1[q] : q < 10 ? 0.0 : q < 20 ? Tx[q - 10] : q < 30 ? Ty[q - 20] : q < 40 ? Tr[q - 30] : 0.0;
#pragma unroll
        for (int lvl = 0; lvl < 5; lvl++) {
            const int half = 32 >> lvl;                   // values a lane holds after this level
            const bool up = (lane >> (4 - lvl)) & 1;
#pragma unroll
            for (int k = 0; k < half; k++) {
                const double keep = up ? v[k + half] : v[k];
                const double send = up ? v[k] : v[k + half];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16 >> lvl);
            }
        }
        if (lane < 20) { s_red[warp][2 * lane] = v[0]; s_red[warp][2 * lane + 1] = v[1]; }
    }
    __syncwarp();
    if (lane == 0) { s_red[warp][9] = (double)wcount; s_red[warp][40] = (double)wcount; s_red[warp][41] = (double)my_first; }
    __syncthreads();
    if (tid < 42) {
        double t = tid == 41 ? 2147483647.0 : 0.0;
#pragma unroll
        for (int q = 0; q < FS_WARPS; q++) t = tid == 41 ? fmin(t, s_red[q][tid]) : t + s_red[q][tid];
        w.part[((size_t)b * FS_SEG + rank) * FS_NQ + tid] = t;
    }
}

__global__ void __launch_bounds__(32) zp_fin_solve_kernel(FinalArgs a, FsWs w) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (w.state[b]) return;
    const bool stamp = w.dbg && b == 0 && lane == 0;
    if (stamp) w.dbg[8] = clock64();
    __shared__ double s_T[FS_NQ];
    __shared__ double s_A[16];
    __shared__ double s_c0[3];
    __shared__ ZpSums s_sums;
    __shared__ ZpControl s_cp;
    __shared__ double s_V[48];
    __shared__ __align__(16) double s_eig[ZP_SYM_DOUBLES + 24];
    for (int q = lane; q < 42; q += 32) {                  // ((p0 + p1) + p2) + p3; the first inlier index is a minimum
        double t = q == 41 ? 2147483647.0 : 0.0;
#pragma unroll
        for (int r = 0; r < FS_SEG; r++) {
            const double v = w.part[((size_t)b * FS_SEG + r) * FS_NQ + q];
            t = q == 41 ? fmin(t, v) : t + v;
        }
        s_T[q] = t;
    }
    __syncwarp();
    const int ni = (int)s_T[40];
    const int first = (int)s_T[41];
    const double* hp = a.hyp_poses + ((size_t)b * a.H + a.rs[4 * b + 2]) * 12;
    if (lane == 0) a.n_inliers[b] = ni;
    if (ni < 4) {       // cannot happen after selection (good > m-1 >= 3) but keep the output defined
        if (lane < 12) a.poses[12 * (size_t)b + lane] = hp[lane];
        if (a.records && lane < 14) a.records[14 * (size_t)b + lane] = lane < 12 ? hp[lane] : lane == 12 ? (double)ni : (double)a.status[b];
        if (lane == 0) w.state[b] = 1;
        return;
    }
    if (stamp) w.dbg[9] = clock64();
    const float* cb = a.corr + (size_t)b * 5 * a.cap;
    const float *pX = cb + 2 * (size_t)a.cap, *pY = cb + 3 * (size_t)a.cap, *pZ = cb + 4 * (size_t)a.cap;
    const double g[3] = {(double)pX[0], (double)pY[0], (double)pZ[0]};
    const double* Kb = a.K + 9 * (size_t)b;
    const ZpCam cam{Kb[0], Kb[4], Kb[2], Kb[5]};
    if (lane == 0) {
        double A[16], c0[3];
        zp_moment_frame(s_T, (double)ni, s_cp, A, c0);
        for (int k = 0; k < 16; k++) s_A[k] = A[k];
        for (int k = 0; k < 3; k++) s_c0[k] = c0[k];
        s_sums.n = (double)ni;
    }
    __syncwarp();
    if (stamp) w.dbg[10] = clock64();
    {
        double* S = (double*)&s_sums;                      // s0[10] | sx[10] | sy[10] | sr[10] | w[12]
        for (int o = lane; o < 52; o += 32) S[o] = zp_moment_sum(o, s_T, s_A, s_c0);
    }
    __syncwarp();
    if (stamp) w.dbg[11] = clock64();
    // ---- 12x12 null space on 16 lanes, the three beta candidates on three lanes
    if (lane < 16) zp_nullspace4<16>(ZpSym12{s_eig}, s_eig + ZP_SYM_DOUBLES, s_eig + ZP_SYM_DOUBLES + 12, s_sums.s0, cam, lane, 0xFFFFu, s_V);
    __syncwarp();
    if (stamp) w.dbg[12] = clock64();
    if (lane < 3) {
        ZpMat V{s_V, 1};
        double L[60], rho[6], af[4], R[9], t[3];
        zp_L_rho(V, s_cp, L, rho);
        ZpHorn hs;
        zp_horn_inputs(s_sums, hs);
        zp_alphas(s_cp, (double)pX[first] - g[0], (double)pY[first] - g[1], (double)pZ[first] - g[2], af);
        const double c0w[3] = {s_c0[0] + g[0], s_c0[1] + g[1], s_c0[2] + g[2]};    // t = pc0 - R c0 wants the world centroid
        const bool ok = zp_candidate(lane, L, rho, V, hs, af, c0w, R, t);
        double* out = w.cand + (size_t)b * FS_CAND + 13 * lane;
        for (int e = 0; e < 9; e++) out[e] = R[e];
        for (int e = 0; e < 3; e++) out[9 + e] = t[e];
        out[12] = ok ? 1.0 : 0.0;
    }
    if (stamp) w.dbg[13] = clock64();
}

__global__ void __launch_bounds__(FS_THREADS, 4) zp_fin_errors_kernel(FinalArgs a, FsWs w) {
    const int b = blockIdx.x / FS_SEG, rank = blockIdx.x % FS_SEG;
    if (w.state[b]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_red[FS_WARPS][3];
    __shared__ int s_last;
    const int n = min(a.counts[b], a.cap);
    const float* cb = a.corr + (size_t)b * 5 * a.cap;
    const float *pu = cb, *pv = cb + a.cap, *pX = cb + 2 * (size_t)a.cap, *pY = cb + 3 * (size_t)a.cap, *pZ = cb + 4 * (size_t)a.cap;
    const double* Kb = a.K + 9 * (size_t)b;
    const ZpCam cam{Kb[0], Kb[4], Kb[2], Kb[5]};
    const double* cd = w.cand + (size_t)b * FS_CAND;
    __shared__ double s_c[3][12];                          // R | t of the candidates (a failed one: identity, t = (0, 0, 1))
    if (tid < 36) {
        const int c = tid / 12, e = tid - 12 * c;
        const bool ok = cd[13 * c + 12] != 0.0;
        s_c[c][e] = ok ? cd[13 * c + e] : (e < 9 ? (e % 4 == 0 ? 1.0 : 0.0) : (e == 11 ? 1.0 : 0.0));
    }
    __syncthreads();
    const int chunk = fs_chunk(n);
    const int list = rank * FS_WARPS + warp;
    const uint16_t* gl = w.idx + (size_t)b * a.cap + (size_t)list * chunk;
    const int wcount = w.wcnt[b * FS_LISTS + list];
    double acc[3] = {0, 0, 0};
    for (int k0 = 0; k0 < wcount; k0 += 128) {            // four points per trip: their indices, then their 20 gathers, up front
        int idx[4];
        float fX[4], fY[4], fZ[4], fu[4], fv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) idx[q] = k0 + 32 * q + lane < wcount ? (int)gl[k0 + 32 * q + lane] : -1;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = idx[q] < 0 ? 0 : idx[q];
            fX[q] = pX[i]; fY[q] = pY[i]; fZ[q] = pZ[i]; fu[q] = pu[i]; fv[q] = pv[i];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (idx[q] < 0) continue;                     // the sums stay in list order: k0 + lane, + 32, + 64, + 96
            const double X = fX[q], Y = fY[q], Z = fZ[q], u = fu[q], v = fv[q];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                // candidates are re-read from shared memory per point (volatile: hoisted out of the loop they would take 72 registers)
                const volatile double* p = s_c[c];
                const double Xc = p[0] * X + p[1] * Y + p[2] * Z + p[9];
                const double Yc = p[3] * X + p[4] * Y + p[5] * Z + p[10];
                const double iz = 1.0 / (p[6] * X + p[7] * Y + p[8] * Z + p[11]);
                const double du = u - (cam.uc + cam.fu * Xc * iz), dv = v - (cam.vc + cam.fv * Yc * iz);
                acc[c] += sqrt(du * du + dv * dv);           // = zp_reproj_dist (epnp::reprojection_error)
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double x = acc[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) s_red[warp][c] = x;
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < 3; c++) {
            double t = 0;
            for (int q = 0; q < FS_WARPS; q++) t += s_red[q][c];
            w.err[((size_t)b * FS_SEG + rank) * 4 + c] = t;
        }
        __threadfence();
        s_last = atomicAdd(&w.done[b], 1) == FS_SEG - 1;
    }
    __syncthreads();
    if (!s_last || tid >= 32) return;
    // ---- the last CTA of the crop: EPnP's choice among the candidates (totals in rank order), pose out
    __threadfence();
    const int ni = a.n_inliers[b];
    int pick = -1;
    double be = 0;
    for (int c = 0; c < 3; c++) {
        if (cd[13 * c + 12] == 0.0) continue;
        double t = 0;
        for (int r = 0; r < FS_SEG; r++) t += ((volatile double*)w.err)[((size_t)b * FS_SEG + r) * 4 + c];
        const double e = t / ni;
        if (!(e == e)) continue;
        if (pick < 0 || e < be) { pick = c; be = e; }
    }
    const double* src = pick < 0 ? a.hyp_poses + ((size_t)b * a.H + a.rs[4 * b + 2]) * 12 : cd + 13 * pick;
    if (lane < 12) a.poses[12 * (size_t)b + lane] = src[lane];
    if (a.records && lane < 14) a.records[14 * (size_t)b + lane] = lane < 12 ? src[lane] : lane == 12 ? (double)ni : (double)a.status[b];
}

static inline size_t fs_align(size_t x) { return (x + 255) & ~(size_t)255; }

int zp_launch_final_split(zp_ctx* ctx, const FinalArgs& a, cudaStream_t st) {
    const int B = a.B;
    const size_t o_idx = 0, o_wc = o_idx + fs_align((size_t)B * a.cap * sizeof(uint16_t)),
                 o_part = o_wc + fs_align((size_t)B * FS_LISTS * 4), o_cand = o_part + fs_align((size_t)B * FS_SEG * FS_NQ * 8),
                 o_err = o_cand + fs_align((size_t)B * FS_CAND * 8), o_done = o_err + fs_align((size_t)B * FS_SEG * 4 * 8),
                 o_state = o_done + fs_align((size_t)B * 4), need = o_state + fs_align((size_t)B * 4);
    if (need > ctx->fws_bytes) {        // growing must not race with work still using the old buffer
        ZP_CUDA(ctx, cudaDeviceSynchronize());
        zp_drop_graphs(ctx);
        if (ctx->fws) cudaFree(ctx->fws);
        ctx->fws = nullptr; ctx->fws_bytes = 0;
        ZP_CUDA(ctx, cudaMalloc(&ctx->fws, need + need / 8));
        ctx->fws_bytes = need + need / 8;
    }
    char* base = (char*)ctx->fws;
    FsWs w;
    w.idx = (uint16_t*)(base + o_idx); w.wcnt = (int32_t*)(base + o_wc); w.part = (double*)(base + o_part);
    w.cand = (double*)(base + o_cand); w.err = (double*)(base + o_err); w.done = (int32_t*)(base + o_done);
    w.state = (int32_t*)(base + o_state);
    w.dbg = (unsigned long long*)ctx->dbg_buf;
    const int chunk_max = ((a.cap + FS_LISTS * 32 - 1) / (FS_LISTS * 32)) * 32;
    const int smem = FS_WARPS * chunk_max * (int)sizeof(uint16_t);
    if (smem > 48 * 1024) ZP_FAIL(ctx, -1, "zp_ransac: cap %d needs %d bytes of shared memory in the final solve", a.cap, smem);
    ZP_TIME_BEGIN(ctx, st);
    zp_fin_moments_kernel<<<B * FS_SEG, FS_THREADS, smem, st>>>(a, w);
    ZP_CHECK_LAUNCH(ctx, "zp_fin_moments_kernel");
    ZP_TIME_BEGIN(ctx, st);
    zp_fin_solve_kernel<<<B, 32, 0, st>>>(a, w);
    ZP_CHECK_LAUNCH(ctx, "zp_fin_solve_kernel");
    ZP_TIME_BEGIN(ctx, st);
    zp_fin_errors_kernel<<<B * FS_SEG, FS_THREADS, 0, st>>>(a, w);
    ZP_CHECK_LAUNCH(ctx, "zp_fin_errors_kernel");
    return 0;
}
