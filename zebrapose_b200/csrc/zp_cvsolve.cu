// Minimal-sample hypotheses of the batched RANSAC, solved EXACTLY as cv2.solvePnP(SOLVEPNP_EPNP) solves them inside
// cv2.solvePnPRansac (/root/reference/zebrapose/binary_code_helper/CNN_output_to_pose.py:155-157): zp_cvepnp.cuh replays
// OpenCV's double-precision operations one by one, so a hypothesis is bit-identical to cv2's and the RANSAC winner is
// cv2's winner.  THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false (zebrapose_b200/_build.py): a contracted multiply-add
// anywhere in the replayed chain changes the null-space basis of a 5-point sample.
//
// The work is bound by the FP64 pipe (16 lanes per scheduler; tools/micro/fp64_lat.cu: a division chain costs 119 cycles
// per link, a square root 88, a warp-wide DADD/DMUL issues every 2 cycles), and a warp instruction costs the same
// whether 1 or 32 lanes are useful -- a first single-kernel version that kept six lanes per hypothesis through all phases
// spent 55 % of its time in phases that use one to three of them (profiles/r2b_cv_phases.txt).  Hence four kernels with
// the lane mapping each stage wants, handing a [field][hypothesis] record through global memory (L2-resident at
// BASELINE's batch sizes):
//   zp_cvs_prep_kernel   thread per hypothesis            stage A: staging ... barycentric coordinates
//   zp_cvs_null_kernel   six lanes per hypothesis         stage B: M^T M, 12x12 Jacobi SVD in the wave-front order, L, rho
//   zp_cvs_cand_kernel   thread per (candidate, hyp.)     stage C: least squares, Gauss-Newton, alignment, error
//   zp_cvs_pick_kernel   thread per hypothesis            stage D: EPnP's choice -> pose, projection matrix
// Hypotheses are solved in waves [h0, h0 + hw) of every crop that has not reached cv2's adaptive stop yet (crop_done).
#include "zp_common.cuh"
#include "zp_cvepnp.cuh"
#include "zp_proj.cuh"
#include <stdlib.h>

constexpr int CVA_THREADS = 64;                // stage A: private data 117 doubles per thread, interleaved over the CTA
constexpr int CVA_PRIV = 24 + 16 + 32 + 12 + 33;
constexpr int CVB_HPW = 5;                     // stage B: hypotheses per warp (5 x 6 lanes; lanes 30 and 31 idle)
constexpr int CVB_WARPS = 2;
constexpr int CVB_HB = 230;                    // doubles per hypothesis: A 12 x 14 = 168 | W 12 | V4 48 (first: al 32 | us 16) | flags 2; even: 16-byte aligned rows
constexpr int CVC_THREADS = 64;                // stage C: 60 doubles of scratch per thread
constexpr int CVC_PRIV = 60;

struct CvsArgs {
    const float* corr; int cap; const int32_t* counts; const double* K; const int32_t* samples;
    int B, H, h0, hw; const int32_t* crop_done; const int32_t* rs; int m; double inv_thr;
    double* rec; int nhp;                      // hand-off records: field f of local hypothesis g at rec[f * nhp + g]
    double* hyp_poses; float* hyp_P; int32_t* hyp_inliers;
    unsigned long long* dbg;                   // profiling aid (zp_debug_buffer): phase stamps of warp 0 of CTA 0 of stage B
};

// local hypothesis index -> (crop, global hypothesis index, must it be solved?)
struct CvsHyp { int b; size_t g; bool live, run; };
__device__ __forceinline__ CvsHyp cvs_hyp(const CvsArgs& a, long long gloc) {
    CvsHyp h;
    h.live = gloc < (long long)a.B * a.hw;
    h.b = h.live ? (int)(gloc / a.hw) : 0;
    h.g = (size_t)h.b * a.H + a.h0 + (h.live ? (int)(gloc - (long long)h.b * a.hw) : 0);
    if (h.live && a.crop_done && a.crop_done[h.b]) h.live = false;
    // hypotheses at or past the crop's current stopping iteration (cv2's niters after the waves so far) are never consulted
    if (h.live && a.rs && (int)(h.g - (size_t)h.b * a.H) >= a.rs[4 * h.b]) h.live = false;
    h.run = false;
    if (h.live) {
        const int n = min(a.counts[h.b], a.cap);
        const int32_t* sidx = a.samples + h.g * a.m;
        bool valid = n >= a.m;
        if (valid)
            for (int j = 0; j < a.m; j++) valid = valid && sidx[j] >= 0 && sidx[j] < n;
        h.run = valid;
    }
    return h;
}

__device__ __forceinline__ CveCam cvs_cam(const CvsArgs& a, int b) {
    const double* Kb = a.K + 9 * (size_t)b;
    return CveCam{Kb[0], Kb[4], Kb[2], Kb[5]};
}

__global__ void __launch_bounds__(CVA_THREADS) zp_cvs_prep_kernel(CvsArgs a) {
    extern __shared__ __align__(16) double s_a[];
    const int tid = threadIdx.x;
    const long long gloc = (long long)blockIdx.x * CVA_THREADS + tid;
    const CvsHyp h = cvs_hyp(a, gloc);
    if (!h.run) return;
    const Dv priv = cve_dv(s_a + tid, CVA_THREADS);
    const Dv pw = priv, us = priv.at(24), al = priv.at(40), cw = priv.at(72), wk = priv.at(84);
    cve_stage_a(a.corr + (size_t)h.b * 5 * a.cap, a.cap, a.samples + h.g * a.m, a.m, cvs_cam(a, h.b), pw, us, al, cw, wk);
    const Dv out = cve_dv(a.rec + gloc, a.nhp);
    for (int k = 0; k < 3 * a.m; k++) out[CVH_PW + k] = pw[k];
    for (int k = 0; k < 2 * a.m; k++) out[CVH_US + k] = us[k];
    for (int k = 0; k < 4 * a.m; k++) out[CVH_AL + k] = al[k];
    for (int k = 0; k < 12; k++) out[CVH_CW + k] = cw[k];
}

template <int MINB>
__global__ void __launch_bounds__(32 * CVB_WARPS, MINB) zp_cvs_null_kernel(CvsArgs a) {
    extern __shared__ __align__(16) double s_b[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = lane / CVE_G, gl = lane - slot * CVE_G;
    const long long gloc = ((long long)blockIdx.x * CVB_WARPS + warp) * CVB_HPW + slot;
    CvsHyp h;
    h.live = false; h.run = false; h.b = 0; h.g = 0;
    if (slot < CVB_HPW) h = cvs_hyp(a, gloc);
    const bool run = h.run;
    if (!__any_sync(0xffffffffu, run)) return;
    double* S = s_b + (size_t)(warp * CVB_HPW + (slot < CVB_HPW ? slot : 0)) * CVB_HB;
    double* A = S; double* W = S + 12 * CVE_RS; double* V4 = W + 12;
    int* flags = (int*)(V4 + 48);
    const Dv rec = cve_dv(a.rec + gloc, a.nhp);
    const bool stamp = a.dbg && blockIdx.x == 0 && tid == 0;
    if (stamp) a.dbg[0] = clock64();
    if (run) {                                 // al | us staged where V4 will be written later
        for (int k = gl; k < 4 * a.m; k += CVE_G) V4[k] = rec[CVH_AL + k];
        for (int k = gl; k < 2 * a.m; k += CVE_G) V4[32 + k] = rec[CVH_US + k];
    }
    __syncwarp();
    {
        const CveCam cam = cvs_cam(a, h.b);
        double sums[13];
        if (run) cve_b_mtm_table(A, gl, V4, V4 + 32, a.m, cam);
        __syncwarp();
        if (run) cve_b_mtm_sums(A, gl, V4, V4 + 32, a.m, cam, sums);
        __syncwarp();
        if (run) cve_b_mtm_store(A, gl, sums);
    }
    CveJ j = cve_j_make(A, CVE_RS, 12, 12, gl, CVE_G, flags, run);
    cve_j_init(j);
    __syncwarp();
    if (stamp) a.dbg[1] = clock64();
    int T = 1;
    for (;; T++) {
        if (!j.done) cve_jstep_a<12, 12>(j, T);
        __syncwarp();
        if (!j.done) cve_jstep_c<12>(j, T);
        if (__all_sync(0xffffffffu, j.done)) break;
    }
    __syncwarp();
    if (stamp) { a.dbg[2] = clock64(); a.dbg[5] = (unsigned long long)T; }
    if (run) cve_b_norms(A, gl, W);
    __syncwarp();
    {
        const bool fast = run ? cve_b_tail(A, gl, W, V4) : true;
        if (run && !fast && gl == 0) cve_b_finish(A, W, V4);      // a zero or non-finite singular value (degenerate sample)
    }
    __syncwarp();
    if (stamp) a.dbg[3] = clock64();
    if (run) {
        cve_b_L_rho(V4, gl, rec.at(CVH_CW), rec.at(CVH_L), rec.at(CVH_RHO));
        for (int k = gl; k < 48; k += CVE_G) rec[CVH_V4 + k] = V4[k];
    }
    if (stamp) a.dbg[4] = clock64();
}

template <int MINB>
__global__ void __launch_bounds__(CVC_THREADS, MINB) zp_cvs_cand_kernel(CvsArgs a) {
    extern __shared__ __align__(16) double s_c[];
    const int tid = threadIdx.x, c = blockIdx.y;
    const long long gloc = (long long)blockIdx.x * CVC_THREADS + tid;
    const CvsHyp h = cvs_hyp(a, gloc);
    if (!h.run) return;
    const Dv rec = cve_dv(a.rec + gloc, a.nhp);
    cve_stage_c(c, a.m, cvs_cam(a, h.b), rec.at(CVH_L), rec.at(CVH_RHO), rec.at(CVH_V4), rec.at(CVH_AL), rec.at(CVH_PW),
                rec.at(CVH_US), cve_dv(s_c + tid, CVC_THREADS), rec.at(CVH_OUT + 13 * c));
}

__global__ void __launch_bounds__(128) zp_cvs_pick_kernel(CvsArgs a) {
    const long long gloc = (long long)blockIdx.x * 128 + threadIdx.x;
    const CvsHyp h = cvs_hyp(a, gloc);
    if (!h.live) return;
    double* out = a.hyp_poses + h.g * 12;
    float4* outP = (float4*)(a.hyp_P + h.g * 24);                // every element twice: (P,P) pairs for FFMA2
    if (a.hyp_inliers) a.hyp_inliers[h.g] = 0;
    if (!h.run) {
        for (int e = 0; e < 12; e++) out[e] = nan("");
        for (int e = 0; e < 6; e++) outP[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const Dv o = cve_dv(a.rec + (size_t)CVH_OUT * a.nhp + gloc, a.nhp);
    const int N = cve_pick(o);
    double pose[12];
    for (int e = 0; e < 12; e++) { pose[e] = o[13 * N + e]; out[e] = pose[e]; }
    float P[12];
    zp_make_P(pose, a.K + 9 * (size_t)h.b, a.inv_thr, P);
    for (int e = 0; e < 6; e++) outP[e] = make_float4(P[2 * e], P[2 * e], P[2 * e + 1], P[2 * e + 1]);
}

int zp_launch_minimal_cv(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                         const int32_t* samples, int B, int H, int h0, int hw, const int32_t* crop_done, const int32_t* rs,
                         int m, float thr_px, double* hyp_poses, float* hyp_P, int32_t* hyp_inliers_to_zero, cudaStream_t st) {
    const int smem_a = CVA_THREADS * CVA_PRIV * (int)sizeof(double);
    const int smem_b = CVB_WARPS * CVB_HPW * CVB_HB * (int)sizeof(double);
    const int smem_c = CVC_THREADS * CVC_PRIV * (int)sizeof(double);
    if (!ctx->cvs_attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_cvs_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_cvs_null_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_cvs_null_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_cvs_null_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_cvs_cand_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_c));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_cvs_cand_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_c));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_cvs_cand_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_c));
        const char* e = getenv("ZP_CVB_MINB");      // tuning aids: resident CTAs per SM of stage B (8 | 10 | 12) and C (4 | 6 | 8)
        ctx->cvb_minb = e ? atoi(e) : 0;
        e = getenv("ZP_CVC_MINB");
        ctx->cvc_minb = e ? atoi(e) : 0;
        ctx->cvs_attr_set = true;
    }
    const long long total = (long long)B * hw;
    const int nhp = (int)((total + 63) / 64 * 64);
    const size_t need = (size_t)CVH_DOUBLES * nhp * sizeof(double);
    if (need > ctx->cvws_bytes) {       // growing must not race with work still using the old buffer
        ZP_CUDA(ctx, cudaDeviceSynchronize());
        zp_drop_graphs(ctx);
        if (ctx->cvws) cudaFree(ctx->cvws);
        ctx->cvws = nullptr; ctx->cvws_bytes = 0;
        ZP_CUDA(ctx, cudaMalloc(&ctx->cvws, need + need / 8));
        ctx->cvws_bytes = need + need / 8;
    }
    CvsArgs a;
    a.corr = corr; a.cap = cap; a.counts = counts; a.K = K; a.samples = samples; a.B = B; a.H = H; a.h0 = h0; a.hw = hw;
    a.crop_done = crop_done; a.rs = rs; a.m = m; a.inv_thr = 1.0 / (double)thr_px; a.rec = (double*)ctx->cvws; a.nhp = nhp;
    a.hyp_poses = hyp_poses; a.hyp_P = hyp_P; a.hyp_inliers = hyp_inliers_to_zero;
    a.dbg = (unsigned long long*)ctx->dbg_buf;
    ZP_TIME_BEGIN(ctx, st);
    zp_cvs_prep_kernel<<<(unsigned)((total + CVA_THREADS - 1) / CVA_THREADS), CVA_THREADS, smem_a, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_cvs_prep_kernel");
    const int per_cta = CVB_WARPS * CVB_HPW;
    ZP_TIME_BEGIN(ctx, st);
    const unsigned grid_b = (unsigned)((total + per_cta - 1) / per_cta);
    const int mb = ctx->cvb_minb ? ctx->cvb_minb : 10;     // measured (profiles/r2e_tune.jsonl): 8 | 10 | 12 -> 1958 | 1840 | 1861 us at 1024 crops
    if (mb >= 12) zp_cvs_null_kernel<12><<<grid_b, 32 * CVB_WARPS, smem_b, st>>>(a);
    else if (mb >= 10) zp_cvs_null_kernel<10><<<grid_b, 32 * CVB_WARPS, smem_b, st>>>(a);
    else zp_cvs_null_kernel<8><<<grid_b, 32 * CVB_WARPS, smem_b, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_cvs_null_kernel");
    ZP_TIME_BEGIN(ctx, st);
    const dim3 grid_c((unsigned)((total + CVC_THREADS - 1) / CVC_THREADS), 3);
    const int mc = ctx->cvc_minb ? ctx->cvc_minb : 6;      // 4 | 6 | 8 -> 811 | 691 | 749 us at 1024 crops, 81 | 87 | 92 at 64
    if (mc >= 8) zp_cvs_cand_kernel<8><<<grid_c, CVC_THREADS, smem_c, st>>>(a);
    else if (mc >= 6) zp_cvs_cand_kernel<6><<<grid_c, CVC_THREADS, smem_c, st>>>(a);
    else zp_cvs_cand_kernel<4><<<grid_c, CVC_THREADS, smem_c, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_cvs_cand_kernel");
    ZP_TIME_BEGIN(ctx, st);
    zp_cvs_pick_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a);
    ZP_CHECK_LAUNCH(ctx, "zp_cvs_pick_kernel");
    return 0;
}
