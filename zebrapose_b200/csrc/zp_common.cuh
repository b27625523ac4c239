// Shared declarations of the zebrapose_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/zebrapose_b200.h"

#define ZP_SM_COUNT_FALLBACK 148
#define ZP_MAX_DEVICES 64
#define ZP_TABLE_ROWS 65536          // rows every device dictionary is padded to (codes are 16 bit)

struct ZpTable {
    float4* pts = nullptr;      // [2^(n_bits-k)] x,y,z,exists  (L2-resident, gathered per masked pixel)
    uint16_t* remap = nullptr;  // [2^(n_bits-k)] Hamming-nearest existing code (identity in ZERO mode)
    int n_bits = 0, ignore_bit = 0, mode = 0;
};

struct ZpModel {
    double* pts = nullptr;      // [V][3] model vertices in mm (float64, as pose_error.add/adi get them)
    int V = 0;
};

// one captured chain (zp_pose_batch_device): every argument the enqueue depends on, compared bytewise
struct ZpGraphKey {
    const void* p[10]; int64_t s[4]; int i[16]; int w[16]; float f; double c; uint64_t seed;
};
struct ZpGraph { ZpGraphKey key; cudaGraphExec_t exec; int kernels; };

struct zp_ctx {
    int device = 0;
    int sm_count = ZP_SM_COUNT_FALLBACK;
    std::string err;
    ZpTable tables[ZP_MAX_OBJECTS];
    const float4** d_table_ptrs = nullptr;   // device array [ZP_MAX_OBJECTS + 1]; the last slot = null_table (bad obj ids land there)
    float4* null_table = nullptr;            // ZP_TABLE_ROWS all-non-existing rows shared by the empty slots
    uint32_t* d_rng = nullptr;               // raw outputs of cv::RNG(0xFFFFFFFFFFFFFFFF), replayed by zp_samples_kernel
    int n_rng = 0;
    int* d_counters = nullptr;               // [2] work-queue ticket + done counter of zp_score_kernel (self re-arming)
    // growable device workspace
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // second workspace for the host-buffer entry (device copies of inputs/outputs)
    void* hws = nullptr;
    size_t hws_bytes = 0;
    // decode workspace of the two-kernel path (codes when the caller does not want them, mask ballot words)
    void* dws = nullptr;
    size_t dws_bytes = 0;
    // evaluation (zp_eval.cu): model vertices per object slot + workspace of zp_pose_errors
    ZpModel models[ZP_MAX_OBJECTS];
    const double** d_model_ptrs = nullptr;   // device array [ZP_MAX_OBJECTS]
    int* d_model_V = nullptr;                // device array [ZP_MAX_OBJECTS]
    void* ews = nullptr;
    size_t ews_bytes = 0;
    // fused network head (zp_head.cu): bf16 weights [32][c_in] (rows >= n_out zero), bias, workspace (codes + mask words)
    void* head_w = nullptr;
    void* head_w32 = nullptr;
    float head_bias[32] = {0};
    int head_n_out = 0, head_c_in = 0;
    void* hdws = nullptr;
    size_t hdws_bytes = 0;
    cudaStream_t own_stream = nullptr;
    // per-kernel timing (zp_set_kernel_timing): CUDA events recorded on the launching stream directly around a launch
    int timing = 0;
    bool t_open = false;
    cudaEvent_t t_ev0 = nullptr, t_ev1 = nullptr;
    cudaStream_t t_stream = nullptr;
    std::map<std::string, std::pair<double, long long>> t_acc;      // kernel name -> (sum of ms, launches)
    int64_t launches = 0;
    int score_groups = 0;                    // 0 auto, else 1 | 2 | 4 warp-groups per scoring CTA (tests / tuning)
    int score_hchunk = 0;                    // 0 auto, -1 never cut, else hypotheses per scoring work item
    void* dbg_buf = nullptr;                 // device buffer for per-CTA timestamps of the decode kernel (zp_debug_buffer)
    int decode_rpc = 0;                      // runs per CTA of the streaming decode kernel (0 = automatic)
    int force_decode_path = 0;               // 0 auto, 1 register-staged cluster kernel, 2 generic kernel (tests)
    // zp_pose_batch_device: correspondence lists between decode and RANSAC, and the captured graphs
    void* gws = nullptr;
    size_t gws_bytes = 0;
    std::vector<ZpGraph> graphs;
    // hand-off records between the stages of the exact minimal solver (zp_cvsolve.cu)
    void* cvws = nullptr;
    size_t cvws_bytes = 0;
    // workspace of the split final solve (zp_finsplit.cu): packed inlier lists, partial sums, candidates
    void* fws = nullptr;
    size_t fws_bytes = 0;
    // RANSAC: minimal solver (ZP_SOLVER_*), wave plan (hypotheses per wave; n_waves = 0: automatic)
    int solver = 0;
    int n_waves = 0;
    int wave_sizes[16] = {0};
    // "function attribute set on this context's device" flags (cudaFuncSetAttribute is per device, a ctx is per device)
    bool cvs_attr_set = false, min_attr_set = false, fin_attr_set = false, fin_form_set = false;
    int min_force = 0, cvb_minb = 0, cvc_minb = 0, fin_force = 0;
    int score_per_sm[3] = {0, 0, 0};
    int rs_no_recount = 0;                   // 1: the adaptive-stop replay decides near-ties on the FP32 counts (tests); 2: exact re-counts without parking (tests)
};

#define ZP_FAIL(ctx, code, ...)                                  \
    do {                                                         \
        char _b[512];                                            \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                   \
        (ctx)->err = _b;                                         \
        return (code);                                           \
    } while (0)

#define ZP_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess)                                                               \
            ZP_FAIL(ctx, -2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// Profiling aid: ZP_TIME_BEGIN right before a kernel launch + the ZP_CHECK_LAUNCH after it bracket exactly that launch
// with two events on its stream (only when zp_set_kernel_timing is on; the end side waits for the kernel).
#define ZP_TIME_BEGIN(ctx, st)                                                               \
    do {                                                                                     \
        if ((ctx)->timing) { cudaEventRecord((ctx)->t_ev0, (st)); (ctx)->t_stream = (st); (ctx)->t_open = true; } \
    } while (0)

#define ZP_CHECK_LAUNCH(ctx, name)                                                           \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess)                                                               \
            ZP_FAIL(ctx, -3, "launch of %s failed: %s", name, cudaGetErrorString(_e));       \
        (ctx)->launches++;                                                                   \
        if ((ctx)->t_open) {                                                                 \
            float _ms = 0.f;                                                                 \
            cudaEventRecord((ctx)->t_ev1, (ctx)->t_stream);                                  \
            cudaEventSynchronize((ctx)->t_ev1);                                              \
            cudaEventElapsedTime(&_ms, (ctx)->t_ev0, (ctx)->t_ev1);                          \
            auto& _a = (ctx)->t_acc[name];                                                   \
            _a.first += _ms; _a.second += 1;                                                 \
            (ctx)->t_open = false;                                                           \
        }                                                                                    \
    } while (0)

// arguments of the final solve (zp_final_cl_kernel in zp_ransac.cu, the split form in zp_finsplit.cu)
struct FinalArgs {
    const float* corr; int cap; const int32_t* counts; const double* K; const double* hyp_poses;
    const int32_t* hyp_inliers; int B, H, m; double conf; int select_mode; float inv_thr; int final_mode;
    double* poses; int32_t* n_inliers; int32_t* status; int32_t* best_idx; uint8_t* inlier_mask;
    const int32_t* rs; int32_t* iters_run;      // per-crop RANSAC state {niters, maxGood, best, iterations run}
    float thr2;                                  // float32(thr_px^2), cv2's comparison value
    double* records;                             // nullable [B,14]: pose | n_inliers | status as doubles (the multi-GPU gather record)
};

// Captured graphs (zp_pose_batch_device) hold raw pointers into the ctx's device workspaces: whenever one of those is
// re-allocated -- always in an eager call, after a cudaDeviceSynchronize -- the graphs are dropped and re-captured on next use.
inline void zp_drop_graphs(zp_ctx* ctx) {
    for (auto& g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
}

int zp_ws_reserve(zp_ctx* ctx, size_t bytes);

// table slot of crop b: ids outside [0, ZP_MAX_OBJECTS) go to the extra all-non-existing slot
__device__ __forceinline__ int zp_obj_slot(const int32_t* obj_ids, int obj_default, int b) {
    const int o = obj_ids ? obj_ids[b] : obj_default;
    return (unsigned)o < (unsigned)ZP_MAX_OBJECTS ? o : ZP_MAX_OBJECTS;
}

// streaming 128-bit load: read once, do not pollute L1
__device__ __forceinline__ uint4 zp_ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// kernels' host launchers (defined in the .cu files)
int zp_launch_decode(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4],
                     int mask_ch, int bit0_ch, int nb, const uint8_t* ext_mask, const double* bbox,
                     const int32_t* obj_ids, int obj_default, uint16_t* codes, float* corr, int cap,
                     int32_t* counts, cudaStream_t st);
