// Minimal-sample hypotheses of the batched RANSAC, solved EXACTLY as cv2.solvePnP(SOLVEPNP_EPNP) solves them inside
// cv2.solvePnPRansac (/root/reference/zebrapose/binary_code_helper/CNN_output_to_pose.py:155-157): zp_cvepnp.cuh replays
// OpenCV's double-precision operations one by one, so a hypothesis is bit-identical to cv2's and the RANSAC winner is
// cv2's winner.  THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false (zebrapose_b200/_build.py): a contracted multiply-add
// anywhere in the replayed chain changes the null-space basis of a 5-point sample.
//
//   zp_minimal_cv_kernel   six lanes per hypothesis, five hypotheses per warp; the hypothesis' matrices live in shared
//                          memory (3.4 KB each).  The Jacobi SVDs (3x3 PCA, 3x3 inverse, 12x12 null space, the three
//                          6xN least-squares problems, three 3x3 alignments) run in the dependency-preserving
//                          wave-front order of zp_cvepnp.cuh: ~145 pair steps instead of ~520 in the serial order.
//                          FP64 latency bound by construction (a step is one dependent chain of 3 divisions, 3 square
//                          roots and a 12-term sequential sum); throughput comes from the 9600+ hypotheses in flight.
//
// Hypotheses are solved in waves [h0, h0 + hw) of every crop that has not reached cv2's adaptive stop yet (crop_done).
#include "zp_common.cuh"
#include "zp_cvepnp.cuh"
#include "zp_proj.cuh"

constexpr int CVS_HPW = 5;                     // hypotheses per warp (5 x 6 lanes; lanes 30 and 31 idle)
constexpr int CVS_WARPS = 2;
constexpr int CVS_THREADS = 32 * CVS_WARPS;
constexpr int CVS_SMEM = CVS_WARPS * CVS_HPW * CVE_HB * (int)sizeof(double);

template <int M, bool HASV>
__device__ __noinline__ void cvs_jrun(CveJ& j) {
    cve_j_init(j);
    __syncwarp();
    for (int T = 1;; T++) {
        if (!j.done) cve_jstep_a<M, HASV>(j, T);
        __syncwarp();
        if (!j.done) cve_jstep_c(j, T);
        if (__all_sync(0xffffffffu, j.done)) break;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(CVS_THREADS, 7)
zp_minimal_cv_kernel(const float* __restrict__ corr, int cap, const int32_t* __restrict__ counts,
                     const double* __restrict__ Kmat, const int32_t* __restrict__ samples, int B, int H, int h0, int hw,
                     const int32_t* __restrict__ crop_done, int m, double inv_thr, double* __restrict__ hyp_poses,
                     float* __restrict__ hyp_P, int32_t* __restrict__ hyp_inliers) {
    extern __shared__ __align__(16) double s_cvs[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = lane / CVE_G, gl = lane - slot * CVE_G;
    const long long gloc = ((long long)blockIdx.x * CVS_WARPS + warp) * CVS_HPW + slot;
    bool live = slot < CVS_HPW && gloc < (long long)B * hw;
    const int b = live ? (int)(gloc / hw) : 0;
    const int h = live ? h0 + (int)(gloc - (long long)b * hw) : 0;
    if (live && crop_done && crop_done[b]) live = false;
    if (!__any_sync(0xffffffffu, live)) return;
    double* S = s_cvs + (size_t)(warp * CVS_HPW + (slot < CVS_HPW ? slot : 0)) * CVE_HB;
    int* flags = (int*)(S + CVE_FLAGS);
    const size_t g = (size_t)b * H + h;
    const int32_t* sidx = samples + g * m;
    const int n = live ? min(counts[b], cap) : 0;
    bool valid = live && n >= m;
    if (valid)
        for (int j = 0; j < m; j++) valid = valid && sidx[j] >= 0 && sidx[j] < n;
    const bool run = valid;
    const double* Kb = Kmat + 9 * (size_t)b;
    CveCam cam{1, 1, 0, 0};
    if (run) { cam.fu = Kb[0]; cam.fv = Kb[4]; cam.uc = Kb[2]; cam.vc = Kb[5]; }
    const float* cb = corr + (size_t)b * 5 * cap;

    if (run) cve_ph0(S, gl, cb, cap, sidx, m, cam);
    CveJ j = run && gl == 0 ? cve_j_make(S + CVE_A3, 3, S + CVE_V3, 3, 3, 3, 0, 1, flags) : cve_j_none();
    cvs_jrun<3, true>(j);
    if (run) cve_ph1(S, gl, m);
    j = run && gl == 0 ? cve_j_make(S + CVE_A3, 3, S + CVE_V3, 3, 3, 3, 0, 1, flags) : cve_j_none();
    cvs_jrun<3, true>(j);
    if (run) cve_ph2(S, gl);
    __syncwarp();
    if (run) cve_ph3(S, gl, m);
    __syncwarp();
    if (run) cve_ph4(S, gl, m, cam);
    j = run ? cve_j_make(S + CVE_A, CVE_RS, nullptr, 0, 12, 12, gl, 6, flags) : cve_j_none();
    cvs_jrun<12, false>(j);
    if (run) cve_ph5(S, gl);
    __syncwarp();
    if (run) cve_ph6(S, gl);
    __syncwarp();
    if (run) cve_ph7(S, gl);
    {
        const int c = cve_lane_cand(gl);
        if (run && c >= 0) {
            const CveCand k = cve_cand(c);
            j = cve_j_make(S + CVE_A + k.at, 6, S + CVE_A + k.vt, k.nc, k.nc, 6, gl - k.lane0, k.nl, flags + 4 * c);
        } else j = cve_j_none();
    }
    cvs_jrun<6, true>(j);
    double betas[4] = {0, 0, 0, 0};
    if (run) cve_ph8(S, gl, m, betas);
    __syncwarp();
    if (run) cve_ph9(S, gl, m, betas);
    {
        const int c = cve_lane_cand(gl);
        if (run && c >= 0 && gl == cve_cand(c).lane0) {
            double* sl = S + CVE_A + 48 * c;
            j = cve_j_make(sl + 24, 3, sl + 33, 3, 3, 3, 0, 1, flags + 4 * c);
        } else j = cve_j_none();
    }
    cvs_jrun<3, true>(j);
    if (run) cve_ph10(S, gl, m, cam);
    __syncwarp();
    if (live && gl == 0) {
        double* out = hyp_poses + g * 12;
        float4* outP = (float4*)(hyp_P + g * 24);                // every element twice: (P,P) pairs for FFMA2
        if (hyp_inliers) hyp_inliers[g] = 0;
        if (!valid) {
            for (int e = 0; e < 12; e++) out[e] = nan("");
            for (int e = 0; e < 6; e++) outP[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const double* o = cve_pick(S);
            double pose[12];
            for (int e = 0; e < 12; e++) { pose[e] = o[e]; out[e] = o[e]; }
            float P[12];
            zp_make_P(pose, Kb, inv_thr, P);
            for (int e = 0; e < 6; e++) outP[e] = make_float4(P[2 * e], P[2 * e], P[2 * e + 1], P[2 * e + 1]);
        }
    }
}

int zp_launch_minimal_cv(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                         const int32_t* samples, int B, int H, int h0, int hw, const int32_t* crop_done, int m,
                         float thr_px, double* hyp_poses, float* hyp_P, int32_t* hyp_inliers_to_zero, cudaStream_t st) {
    if (!ctx->cvs_attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_minimal_cv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CVS_SMEM));
        ctx->cvs_attr_set = true;
    }
    const long long total = (long long)B * hw;
    const int per_cta = CVS_WARPS * CVS_HPW;
    const int grid = (int)((total + per_cta - 1) / per_cta);
    ZP_TIME_BEGIN(ctx, st);
    zp_minimal_cv_kernel<<<grid, CVS_THREADS, CVS_SMEM, st>>>(corr, cap, counts, K, samples, B, H, h0, hw, crop_done, m,
                                                               1.0 / (double)thr_px, hyp_poses, hyp_P, hyp_inliers_to_zero);
    ZP_CHECK_LAUNCH(ctx, "zp_minimal_cv_kernel");
    return 0;
}
