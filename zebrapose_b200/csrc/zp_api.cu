// C ABI of libzebrapose_b200.so (declared in include/zebrapose_b200.h): context, dictionary tables, orchestration.
#include <algorithm>
#include <cmath>
#include "zp_common.cuh"
#include <nvtx3/nvToolsExt.h>      // header-only: ranges show up in Nsight Systems / Compute timelines, no-ops otherwise

int zp_launch_samples(zp_ctx*, const int32_t*, int, int, int, int, int, uint64_t, int32_t*, cudaStream_t);
int zp_launch_minimal(zp_ctx*, const float*, int, const int32_t*, const double*, const int32_t*, int, int, int, int,
                      const int32_t*, const int32_t*, int, float, double*, float*, int32_t*, cudaStream_t);
int zp_launch_poses_to_P(zp_ctx*, const double*, const double*, int, int, float, float*, cudaStream_t);
int zp_launch_score(zp_ctx*, const float*, int, const int32_t*, const float*, int, int, int, int, const int32_t*,
                    const int32_t*, float, int32_t*, bool, cudaStream_t);
int zp_launch_final(zp_ctx*, const float*, int, const int32_t*, const double*, const double*, const int32_t*, int, int, int,
                    double, int, float, int, double*, int32_t*, int32_t*, int32_t*, uint8_t*, const int32_t*, int32_t*,
                    double*, cudaStream_t);
int zp_launch_rs_init(zp_ctx*, const int32_t*, int, int, int, int32_t*, int32_t*, int32_t*, cudaStream_t);
int zp_launch_rs_replay(zp_ctx*, const int32_t*, int, const int32_t*, int, int, int, int, int, double, int, int32_t*,
                        int32_t*, const float*, const double*, const double*, float, cudaStream_t);
int zp_launch_fma_probe(zp_ctx*, int, int, double*);
int zp_launch_dfma_probe(zp_ctx*, int, double*);
int zp_read_debug_clocks(long long*);
int zp_launch_remap_pixels(zp_ctx*, const int64_t*, int64_t, const double*, int, int64_t*, cudaStream_t);
int zp_launch_decode_ce(zp_ctx*, const void*, int, int, int, const int64_t*, int, int, int, int, const uint8_t*, const double*,
                        const int32_t*, int, uint16_t*, float*, int, int32_t*, cudaStream_t);
int zp_launch_codes_to_ids(zp_ctx*, const double*, int64_t, int, int, double*, cudaStream_t);

static thread_local std::string g_err;

struct ZpRange {                      // NVTX range covering the enqueue of one public entry / one stage of the chain
    explicit ZpRange(const char* name) { nvtxRangePushA(name); }
    ~ZpRange() { nvtxRangePop(); }
};

int zp_ws_reserve(zp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return 0;
    // growing the workspace must not race with work still using the old one
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    zp_drop_graphs(ctx);
    if (ctx->ws) cudaFree(ctx->ws);
    ctx->ws = nullptr; ctx->ws_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    ZP_CUDA(ctx, cudaMalloc(&ctx->ws, want));
    ctx->ws_bytes = want;
    return 0;
}

static int hws_reserve(zp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->hws_bytes) return 0;
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->hws) cudaFree(ctx->hws);
    ctx->hws = nullptr; ctx->hws_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    ZP_CUDA(ctx, cudaMalloc(&ctx->hws, want));
    ctx->hws_bytes = want;
    return 0;
}

extern "C" {

int zp_version(void) { return 100; }

int zp_create(zp_ctx** out, int device) {
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_err = "no CUDA device: zebrapose_b200 has no CPU fallback"; return -2; }
    if (device < 0 || device >= ndev) { g_err = "bad device index"; return -1; }
    zp_ctx* ctx = new zp_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; g_err = "cudaSetDevice failed"; return -2; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; g_err = "cudaGetDeviceProperties failed"; return -2; }
    if (prop.major < 9) { delete ctx; g_err = "device too old: built for sm_100a (thread-block clusters, TMA)"; return -2; }
    ctx->sm_count = prop.multiProcessorCount;
    // Every slot of the device pointer array -- plus one extra slot that out-of-range ids are mapped to -- always points at
    // ZP_TABLE_ROWS entries: empty slots share an all-non-existing table, shorter dictionaries are zero padded.  A crop
    // whose obj id is wrong therefore decodes against "no code exists" ((0,0,0) points, RANSAC status NO_MODEL) instead of
    // reading through a null or short pointer.
    {
        std::vector<float4*> ptrs(ZP_MAX_OBJECTS + 1);
        if (cudaMalloc((void**)&ctx->null_table, ZP_TABLE_ROWS * sizeof(float4)) != cudaSuccess ||
            cudaMemset(ctx->null_table, 0, ZP_TABLE_ROWS * sizeof(float4)) != cudaSuccess ||
            cudaMalloc((void**)&ctx->d_table_ptrs, (ZP_MAX_OBJECTS + 1) * sizeof(float4*)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx; g_err = "context allocation failed"; return -2;
        }
        for (auto& q : ptrs) q = ctx->null_table;
        if (cudaMemcpy((void*)ctx->d_table_ptrs, ptrs.data(), ptrs.size() * sizeof(float4*), cudaMemcpyHostToDevice) != cudaSuccess) {
            delete ctx; g_err = "context allocation failed"; return -2;
        }
    }
    {   // raw stream of cv::RNG(0xFFFFFFFFFFFFFFFF) (multiply-with-carry), replayed by zp_samples_kernel
        const int n = 16384;
        std::vector<uint32_t> tab(n);
        uint64_t state = 0xFFFFFFFFFFFFFFFFull;
        for (int i = 0; i < n; i++) {
            state = (uint64_t)(uint32_t)state * 4164903690ull + (state >> 32);
            tab[i] = (uint32_t)state;
        }
        if (cudaMalloc((void**)&ctx->d_rng, n * sizeof(uint32_t)) != cudaSuccess ||
            cudaMemcpy(ctx->d_rng, tab.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            delete ctx; g_err = "context allocation failed"; return -2;
        }
        ctx->n_rng = n;
        if (cudaMalloc((void**)&ctx->d_counters, 2 * sizeof(int)) != cudaSuccess ||
            cudaMemset(ctx->d_counters, 0, 2 * sizeof(int)) != cudaSuccess) {
            delete ctx; g_err = "context allocation failed"; return -2;
        }
    }
    *out = ctx;
    return 0;
}

void zp_destroy(zp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& t : ctx->tables) { if (t.pts) cudaFree(t.pts); if (t.remap) cudaFree(t.remap); }
    if (ctx->d_table_ptrs) cudaFree((void*)ctx->d_table_ptrs);
    if (ctx->null_table) cudaFree(ctx->null_table);
    if (ctx->d_rng) cudaFree(ctx->d_rng);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->hws) cudaFree(ctx->hws);
    if (ctx->cvws) cudaFree(ctx->cvws);
    if (ctx->fws) cudaFree(ctx->fws);
    for (auto& g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->gws) cudaFree(ctx->gws);
    if (ctx->dws) cudaFree(ctx->dws);
    for (auto& m : ctx->models) if (m.pts) cudaFree(m.pts);
    if (ctx->d_model_ptrs) cudaFree((void*)ctx->d_model_ptrs);
    if (ctx->d_model_V) cudaFree(ctx->d_model_V);
    if (ctx->ews) cudaFree(ctx->ews);
    if (ctx->head_w) cudaFree(ctx->head_w);
    if (ctx->head_w32) cudaFree(ctx->head_w32);
    if (ctx->hdws) cudaFree(ctx->hdws);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->t_ev0) cudaEventDestroy(ctx->t_ev0);
    if (ctx->t_ev1) cudaEventDestroy(ctx->t_ev1);
    delete ctx;
}

const char* zp_last_error(zp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int64_t zp_launch_count(zp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zp_set_score_groups(zp_ctx* ctx, int groups, int hyp_chunk) {
    if (!ctx) return -1;
    if (groups != 0 && groups != 1 && groups != 2 && groups != 4) ZP_FAIL(ctx, -1, "zp_set_score_groups: groups must be 0, 1, 2 or 4");
    if (hyp_chunk < -1 || hyp_chunk > ZP_MAX_HYPOTHESES) ZP_FAIL(ctx, -1, "zp_set_score_groups: bad hyp_chunk %d", hyp_chunk);
    ctx->score_groups = groups;
    ctx->score_hchunk = hyp_chunk;
    return 0;
}

int zp_set_solver(zp_ctx* ctx, int solver) {
    if (!ctx) return -1;
    if (solver != ZP_SOLVER_CV2 && solver != ZP_SOLVER_FAST) ZP_FAIL(ctx, -1, "zp_set_solver: bad solver %d", solver);
    ctx->solver = solver;
    return 0;
}

int zp_set_exact_ties(zp_ctx* ctx, int on) {
    if (!ctx) return -1;
    if (on < 0 || on > 2) ZP_FAIL(ctx, -1, "zp_set_exact_ties: 0 (off), 1 (on) or 2 (on, no parking)");
    ctx->rs_no_recount = on == 1 ? 0 : on == 0 ? 1 : 2;
    return 0;
}

int zp_set_final_form(zp_ctx* ctx, int form) {
    if (!ctx) return -1;
    if (form != 0 && form != 1 && form != 2 && form != 4) ZP_FAIL(ctx, -1, "zp_set_final_form: 0 (automatic), 1 (one CTA per crop), 2 (split: three kernels) or 4 (cluster of 4)");
    ctx->fin_force = form;
    ctx->fin_form_set = true;
    return 0;
}

int zp_set_waves(zp_ctx* ctx, int n, const int32_t* sizes) {
    if (!ctx) return -1;
    if (n < 0 || n > 16 || (n > 0 && !sizes)) ZP_FAIL(ctx, -1, "zp_set_waves: 0 <= n <= 16 sizes");
    for (int i = 0; i < n; i++) if (sizes[i] < 1) ZP_FAIL(ctx, -1, "zp_set_waves: wave size %d", sizes[i]);
    ctx->n_waves = n;
    for (int i = 0; i < n; i++) ctx->wave_sizes[i] = sizes[i];
    return 0;
}

int zp_set_kernel_timing(zp_ctx* ctx, int on) {
    if (!ctx) return -1;
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (on && !ctx->t_ev0) {
        ZP_CUDA(ctx, cudaEventCreate(&ctx->t_ev0));
        ZP_CUDA(ctx, cudaEventCreate(&ctx->t_ev1));
    }
    ctx->timing = on ? 1 : 0;
    ctx->t_open = false;
    if (on) ctx->t_acc.clear();
    return 0;
}

int zp_kernel_time(zp_ctx* ctx, const char* kernel_name, double* ms_sum, int64_t* launches) {
    if (!ctx || !kernel_name) return -1;
    auto it = ctx->t_acc.find(kernel_name);
    if (ms_sum) *ms_sum = it == ctx->t_acc.end() ? 0.0 : it->second.first;
    if (launches) *launches = it == ctx->t_acc.end() ? 0 : it->second.second;
    return 0;
}

int zp_debug_buffer(zp_ctx* ctx, void* dev_u64) {
    if (!ctx) return -1;
    ctx->dbg_buf = dev_u64;
    return 0;
}

int zp_set_decode_path(zp_ctx* ctx, int path) {
    if (!ctx) return -1;
    if (path < 0 || (path > 6 && path < 100) || path > 116) ZP_FAIL(ctx, -1, "zp_set_decode_path: path must be 0..6 or 100+runs_per_cta");
    if (path >= 100) { ctx->force_decode_path = 0; ctx->decode_rpc = path - 100; }      // 100 = automatic again
    else { ctx->force_decode_path = path; }
    return 0;
}

int zp_upload_tables(zp_ctx* ctx, int obj_id, const double* pts, int n_bits, int ignore_bit, int mode) {
    if (!ctx) return -1;
    if (obj_id < 0 || obj_id >= ZP_MAX_OBJECTS) ZP_FAIL(ctx, -1, "obj_id %d out of range [0,%d)", obj_id, ZP_MAX_OBJECTS);
    if (n_bits < 1 || n_bits > 16 || ignore_bit < 0 || ignore_bit >= n_bits) ZP_FAIL(ctx, -1, "bad n_bits/ignore_bit %d/%d", n_bits, ignore_bit);
    if (mode != ZP_NONEXIST_ZERO && mode != ZP_NONEXIST_HAMMING) ZP_FAIL(ctx, -1, "bad nonexist mode %d", mode);
    if (!pts) ZP_FAIL(ctx, -1, "null table");
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int k = ignore_bit, nb = n_bits - k;
    const size_t np = (size_t)1 << nb, nc = (size_t)1 << k;
    std::vector<double> val(np * 3);
    std::vector<uint8_t> exists(np);
    for (size_t p = 0; p < np; p++) {
        double acc[3] = {0, 0, 0};
        int cnt = 0;
        for (size_t j = 0; j < nc; j++) {             // ascending child id, sequential float64 adds (generate_new_dict.py:25-31)
            const double* c = pts + (p * nc + j) * 3;
            bool nanrow = std::isnan(c[0]) || std::isnan(c[1]) || std::isnan(c[2]);
            if (mode == ZP_NONEXIST_HAMMING && nanrow) continue;
            acc[0] = acc[0] + c[0]; acc[1] = acc[1] + c[1]; acc[2] = acc[2] + c[2];
            cnt++;
        }
        double div = mode == ZP_NONEXIST_HAMMING ? (double)cnt : (double)nc;
        if (k == 0) { for (int e = 0; e < 3; e++) val[p * 3 + e] = mode == ZP_NONEXIST_HAMMING && cnt == 0 ? NAN : pts[p * 3 + e]; }
        else { for (int e = 0; e < 3; e++) val[p * 3 + e] = cnt ? acc[e] / div : NAN; }
        exists[p] = !(std::isnan(val[p * 3]) || std::isnan(val[p * 3 + 1]) || std::isnan(val[p * 3 + 2]));
    }
    std::vector<uint16_t> remap(np);
    for (size_t p = 0; p < np; p++) remap[p] = (uint16_t)p;
    if (mode == ZP_NONEXIST_HAMMING) {
        // xor patterns ordered by (popcount, value): fewest flipped bits, then the least significant (finest) flips
        std::vector<uint32_t> pat(np);
        for (size_t x = 0; x < np; x++) pat[x] = (uint32_t)x;
        std::sort(pat.begin(), pat.end(), [](uint32_t a, uint32_t b) {
            int pa = __builtin_popcount(a), pb = __builtin_popcount(b);
            return pa != pb ? pa < pb : a < b;
        });
        bool any = false;
        for (size_t p = 0; p < np; p++) any = any || exists[p];
        for (size_t p = 0; p < np && any; p++) {
            if (exists[p]) continue;
            for (size_t q = 1; q < np; q++) {
                size_t e = p ^ pat[q];
                if (exists[e]) { remap[p] = (uint16_t)e; break; }
            }
        }
        if (!any) for (size_t p = 0; p < np; p++) remap[p] = 0;
    }
    std::vector<float> tab(np * 4);
    for (size_t p = 0; p < np; p++) {
        size_t s = remap[p];
        bool ex = exists[s];
        // non-existing rows: 3D point stays (0,0,0) and the pixel is kept (CNN_output_to_pose.py:58-62)
        tab[p * 4 + 0] = ex ? (float)val[s * 3 + 0] : 0.f;
        tab[p * 4 + 1] = ex ? (float)val[s * 3 + 1] : 0.f;
        tab[p * 4 + 2] = ex ? (float)val[s * 3 + 2] : 0.f;
        tab[p * 4 + 3] = exists[p] ? 1.f : 0.f;
    }
    ZpTable& t = ctx->tables[obj_id];
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    if (t.pts) cudaFree(t.pts);
    if (t.remap) cudaFree(t.remap);
    t.pts = nullptr; t.remap = nullptr;
    ZP_CUDA(ctx, cudaMalloc((void**)&t.pts, ZP_TABLE_ROWS * sizeof(float4)));        // zero padded: any 16-bit code is in bounds
    ZP_CUDA(ctx, cudaMalloc((void**)&t.remap, ZP_TABLE_ROWS * sizeof(uint16_t)));
    ZP_CUDA(ctx, cudaMemset(t.pts, 0, ZP_TABLE_ROWS * sizeof(float4)));
    ZP_CUDA(ctx, cudaMemset(t.remap, 0, ZP_TABLE_ROWS * sizeof(uint16_t)));
    ZP_CUDA(ctx, cudaMemcpy(t.pts, tab.data(), np * sizeof(float4), cudaMemcpyHostToDevice));
    ZP_CUDA(ctx, cudaMemcpy(t.remap, remap.data(), np * sizeof(uint16_t), cudaMemcpyHostToDevice));
    ZP_CUDA(ctx, cudaMemcpy((void*)(ctx->d_table_ptrs + obj_id), &t.pts, sizeof(float4*), cudaMemcpyHostToDevice));
    t.n_bits = n_bits; t.ignore_bit = k; t.mode = mode;
    return 0;
}

int zp_download_tables(zp_ctx* ctx, int obj_id, float* pts_out, uint16_t* remap_out) {
    if (!ctx) return -1;
    if (obj_id < 0 || obj_id >= ZP_MAX_OBJECTS || !ctx->tables[obj_id].pts) ZP_FAIL(ctx, -1, "no table in slot %d", obj_id);
    const ZpTable& t = ctx->tables[obj_id];
    size_t np = (size_t)1 << (t.n_bits - t.ignore_bit);
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    if (pts_out) ZP_CUDA(ctx, cudaMemcpy(pts_out, t.pts, np * sizeof(float4), cudaMemcpyDeviceToHost));
    if (remap_out) ZP_CUDA(ctx, cudaMemcpy(remap_out, t.remap, np * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    return 0;
}

int zp_decode(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4], int mask_ch,
              int bit0_ch, int n_bits, int ignore_bit, const uint8_t* ext_mask, const double* bbox,
              const int32_t* obj_ids, int obj_default, uint16_t* codes, float* corr, int cap, int32_t* counts,
              void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!logits || !bbox || !corr || !counts || !strides) ZP_FAIL(ctx, -1, "zp_decode: null argument");
    if (dtype != ZP_DTYPE_F32 && dtype != ZP_DTYPE_BF16) ZP_FAIL(ctx, -1, "zp_decode: dtype %d not supported", dtype);
    if (B < 0 || S <= 0 || S > 4096 || cap <= 0) ZP_FAIL(ctx, -1, "zp_decode: bad B/S/cap %d/%d/%d", B, S, cap);
    if (n_bits < 1 || n_bits > 16 || ignore_bit < 0 || ignore_bit >= n_bits) ZP_FAIL(ctx, -1, "zp_decode: bad n_bits/ignore_bit");
    if (bit0_ch < 0 || (mask_ch < 0 && !ext_mask)) ZP_FAIL(ctx, -1, "zp_decode: negative channel index (mask_ch %d, bit0_ch %d)", mask_ch, bit0_ch);
    if (!obj_ids) {
        if (obj_default < 0 || obj_default >= ZP_MAX_OBJECTS || !ctx->tables[obj_default].pts)
            ZP_FAIL(ctx, -1, "zp_decode: no dictionary uploaded for object slot %d", obj_default);
        const ZpTable& t = ctx->tables[obj_default];
        if (t.n_bits != n_bits || t.ignore_bit != ignore_bit)
            ZP_FAIL(ctx, -1, "zp_decode: slot %d holds a %d-bit/ignore %d table, call asks %d/%d", obj_default, t.n_bits, t.ignore_bit, n_bits, ignore_bit);
    }
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    ZpRange range("zp_decode");
    return zp_launch_decode(ctx, logits, dtype, B, S, strides, mask_ch, bit0_ch, n_bits - ignore_bit, ext_mask, bbox,
                            obj_ids, obj_default, codes, corr, cap, counts, (cudaStream_t)stream);
}

int zp_decode_ce(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4], int mask_ch,
                 int digit0_ch, int base, int n_digits, const uint8_t* ext_mask, const double* bbox,
                 const int32_t* obj_ids, int obj_default, uint16_t* codes, float* corr, int cap, int32_t* counts,
                 void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!logits || !bbox || !corr || !counts || !strides) ZP_FAIL(ctx, -1, "zp_decode_ce: null argument");
    if (dtype != ZP_DTYPE_F32 && dtype != ZP_DTYPE_BF16) ZP_FAIL(ctx, -1, "zp_decode_ce: dtype %d not supported", dtype);
    if (B < 0 || S <= 0 || S > 1024 || S % 4 != 0 || cap <= 0) ZP_FAIL(ctx, -1, "zp_decode_ce: bad B/S/cap %d/%d/%d (S a multiple of 4)", B, S, cap);
    if (base < 2 || base > 256 || n_digits < 1 || n_digits > 16) ZP_FAIL(ctx, -1, "zp_decode_ce: bad base/n_digits %d/%d", base, n_digits);
    double classes = 1;
    for (int d = 0; d < n_digits; d++) classes *= base;
    if (classes > 65536.0) ZP_FAIL(ctx, -1, "zp_decode_ce: base^n_digits = %.0f exceeds the 16-bit class ids of the path", classes);
    for (int o = 0; o < ZP_MAX_OBJECTS; o++) {
        if (obj_ids ? !ctx->tables[o].pts : o != obj_default) continue;
        const ZpTable& t = ctx->tables[o];
        if (!t.pts) ZP_FAIL(ctx, -1, "zp_decode_ce: no dictionary uploaded for object slot %d", o);
        if (t.ignore_bit != 0 || (double)((size_t)1 << t.n_bits) < classes)
            ZP_FAIL(ctx, -1, "zp_decode_ce: slot %d holds a %d-bit/ignore %d table, %0.f classes need ignore_bit 0 and 2^n_bits >= classes", o, t.n_bits, t.ignore_bit, classes);
    }
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    return zp_launch_decode_ce(ctx, logits, dtype, B, S, strides, mask_ch, digit0_ch, base, n_digits, ext_mask, bbox, obj_ids,
                               obj_default, codes, corr, cap, counts, (cudaStream_t)stream);
}

int zp_remap_pixels(zp_ctx* ctx, const int64_t* px, int64_t N, const double* h_bbox, int S, int64_t* out, void* stream) {
    if (!ctx) return -1;
    if (N < 0 || S <= 0 || !h_bbox || (N > 0 && (!px || !out))) ZP_FAIL(ctx, -1, "zp_remap_pixels: bad argument");
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    return zp_launch_remap_pixels(ctx, px, N, h_bbox, S, out, (cudaStream_t)stream);
}

int zp_codes_to_ids(zp_ctx* ctx, const double* bits, int64_t N, int L, int base, double* ids, void* stream) {
    if (!ctx) return -1;
    if (N < 0 || L < 1 || L > 64 || base < 2 || (N > 0 && (!bits || !ids))) ZP_FAIL(ctx, -1, "zp_codes_to_ids: bad argument");
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    return zp_launch_codes_to_ids(ctx, bits, N, L, base, ids, (cudaStream_t)stream);
}

int zp_make_samples(zp_ctx* ctx, const int32_t* counts, int cap, int B, int H, int m, int mode, uint64_t seed,
                    int32_t* samples, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!counts || !samples) ZP_FAIL(ctx, -1, "zp_make_samples: null argument");
    if (m < 4 || m > 8 || H < 1 || H > ZP_MAX_HYPOTHESES) ZP_FAIL(ctx, -1, "zp_make_samples: bad m/H %d/%d", m, H);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    return zp_launch_samples(ctx, counts, cap, B, H, m, mode, seed, samples, (cudaStream_t)stream);
}

int zp_solve_minimal(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K,
                     const int32_t* samples, int B, int H, int m, double* hyp_poses, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!corr || !counts || !K || !samples || !hyp_poses) ZP_FAIL(ctx, -1, "zp_solve_minimal: null argument");
    if (m < 4 || m > 8 || H < 1 || H > ZP_MAX_HYPOTHESES) ZP_FAIL(ctx, -1, "zp_solve_minimal: bad m/H %d/%d", m, H);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (zp_ws_reserve(ctx, (size_t)B * H * 24 * sizeof(float))) return -2;
    return zp_launch_minimal(ctx, corr, cap, counts, K, samples, B, H, 0, H, nullptr, nullptr, m, 2.0f, hyp_poses,
                             (float*)ctx->ws, nullptr, (cudaStream_t)stream);
}

static int check_corr(zp_ctx* ctx, const float* corr, int cap, const char* who) {
    if (cap <= 0 || cap % 4 != 0) ZP_FAIL(ctx, -1, "%s: cap must be a positive multiple of 4 (TMA 16-byte granules), got %d", who, cap);
    if ((uintptr_t)corr % 16 != 0) ZP_FAIL(ctx, -1, "%s: corr must be 16-byte aligned", who);
    return 0;
}

int zp_score(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K, const double* hyp_poses,
             int B, int H, float thr_px, int32_t* hyp_inliers, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!corr || !counts || !K || !hyp_poses || !hyp_inliers) ZP_FAIL(ctx, -1, "zp_score: null argument");
    if (H < 1 || H > ZP_MAX_HYPOTHESES) ZP_FAIL(ctx, -1, "zp_score: bad H %d", H);
    if (check_corr(ctx, corr, cap, "zp_score")) return -1;
    if (!(thr_px > 0)) ZP_FAIL(ctx, -1, "zp_score: thr_px must be > 0");
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (zp_ws_reserve(ctx, (size_t)B * H * 24 * sizeof(float))) return -2;
    if (int r = zp_launch_poses_to_P(ctx, hyp_poses, K, B, H, thr_px, (float*)ctx->ws, (cudaStream_t)stream)) return r;
    return zp_launch_score(ctx, corr, cap, counts, (const float*)ctx->ws, B, H, 0, H, nullptr, nullptr, thr_px, hyp_inliers,
                           false, (cudaStream_t)stream);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// hypotheses per wave: the caller's plan (zp_set_waves), or automatic.  cv2 stops after 24-46 iterations on crops with
// ~70 % inliers and after ~137 at 50 %.  A wave costs a fixed latency (six dependent launches) whatever its size, so a batch
// that does not fill the GPU (<= 128 crops) takes ONE wave; larger batches probe with 32 hypotheses and then run, per crop,
// only the hypotheses below the niters that probe left (the kernels skip the rest hypothesis by hypothesis).
static int wave_size(const zp_ctx* ctx, int w, int B) {
    if (ctx->n_waves > 0) return ctx->wave_sizes[w < ctx->n_waves ? w : ctx->n_waves - 1];
    if (B <= 128) return ZP_MAX_HYPOTHESES;            // latency-bound: one wave of everything
    return w == 0 ? 32 : ZP_MAX_HYPOTHESES;           // throughput-bound: probe 32, then only what each crop's niters still asks for
}

static int ransac_impl(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K, const int32_t* samples,
                       int B, int H, int m, float thr_px, double confidence, int sampler, uint64_t seed, int select_mode,
                       int final_mode, double* hyp_poses, int32_t* hyp_inliers, int32_t* best_idx, int32_t* iters_run,
                       uint8_t* inlier_mask, double* poses, int32_t* n_inliers, int32_t* status, double* records, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!corr || !counts || !K || !poses || !n_inliers || !status) ZP_FAIL(ctx, -1, "zp_ransac: null argument");
    if (m < 4 || m > 8 || H < 1 || H > ZP_MAX_HYPOTHESES) ZP_FAIL(ctx, -1, "zp_ransac: bad m/H %d/%d", m, H);
    if (cap > 65536) ZP_FAIL(ctx, -1, "zp_ransac: cap %d > 65536 not supported", cap);
    if (check_corr(ctx, corr, cap, "zp_ransac")) return -1;
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    ZpRange range("zp_ransac");
    cudaStream_t st = (cudaStream_t)stream;
    if (!(thr_px > 0)) ZP_FAIL(ctx, -1, "zp_ransac: thr_px must be > 0");
    // workspace: hyp_P | RANSAC state | done flags | samples | hyp_poses | hyp_inliers (the last three only if the caller
    // did not supply them)
    size_t o_P = 0, o_rs = o_P + align256((size_t)B * H * 24 * 4), o_dn = o_rs + align256((size_t)B * 4 * 4);
    size_t o_s = o_dn + align256((size_t)B * 4 * 5);        // done flags [B] | near-tie state [B][4]
    size_t o_p = o_s + (samples ? 0 : align256((size_t)B * H * m * 4));
    size_t o_i = o_p + (hyp_poses ? 0 : align256((size_t)B * H * 12 * 8));
    size_t total = o_i + (hyp_inliers ? 0 : align256((size_t)B * H * 4));
    if (total && zp_ws_reserve(ctx, total)) return -2;
    char* ws = (char*)ctx->ws;
    int32_t* d_samples = nullptr;
    if (!samples) {
        d_samples = (int32_t*)(ws + o_s);
        ZpRange rs("samples");
        if (int r = zp_launch_samples(ctx, counts, cap, B, H, m, sampler, seed, d_samples, st)) return r;
    }
    double* d_hp = hyp_poses ? hyp_poses : (double*)(ws + o_p);
    int32_t* d_hi = hyp_inliers ? hyp_inliers : (int32_t*)(ws + o_i);
    float* d_P = (float*)(ws + o_P);
    int32_t* d_rs = (int32_t*)(ws + o_rs);
    int32_t* d_done = (int32_t*)(ws + o_dn);
    // exported hypothesis lists are diagnostic: all H hypotheses of every crop in one wave
    const bool full = hyp_poses || hyp_inliers || select_mode != ZP_SELECT_CV2_REPLAY;
    if (int r = zp_launch_rs_init(ctx, counts, cap, B, H, d_rs, d_done, nullptr, st)) return r;
    for (int h0 = 0, w = 0; h0 < H; w++) {
        int hw = full ? H : wave_size(ctx, w, B);
        if (hw > H - h0) hw = H - h0;
        ZpRange rw("wave: minimal solver + scoring + adaptive-stop replay");
        // the minimal solver zeroes the inlier counters of its hypotheses: the scoring launch follows without a memset node
        const int32_t* d_lim = select_mode == ZP_SELECT_CV2_REPLAY && !full ? d_rs : nullptr;
        if (int r = zp_launch_minimal(ctx, corr, cap, counts, K, samples ? samples : d_samples, B, H, h0, hw, d_done, d_lim, m,
                                      thr_px, d_hp, d_P, d_hi, st)) return r;
        if (int r = zp_launch_score(ctx, corr, cap, counts, d_P, B, H, h0, hw, d_done, d_lim, thr_px, d_hi, true, st)) return r;
        if (int r = zp_launch_rs_replay(ctx, counts, cap, d_hi, B, H, h0, h0 + hw, m, confidence, select_mode, d_rs, d_done, corr, K, d_hp,
                                        thr_px, st))
            return r;
        h0 += hw;
    }
    ZpRange rf("final solve on the inliers");
    return zp_launch_final(ctx, corr, cap, counts, K, d_hp, d_hi, B, H, m, confidence, select_mode, thr_px, final_mode,
                           poses, n_inliers, status, best_idx, inlier_mask, d_rs, iters_run, records, st);
}

int zp_ransac(zp_ctx* ctx, const float* corr, int cap, const int32_t* counts, const double* K, const int32_t* samples,
              int B, int H, int m, float thr_px, double confidence, int sampler, uint64_t seed, int select_mode,
              int final_mode, double* hyp_poses, int32_t* hyp_inliers, int32_t* best_idx, int32_t* iters_run,
              uint8_t* inlier_mask, double* poses, int32_t* n_inliers, int32_t* status, void* stream) {
    return ransac_impl(ctx, corr, cap, counts, K, samples, B, H, m, thr_px, confidence, sampler, seed, select_mode, final_mode,
                       hyp_poses, hyp_inliers, best_idx, iters_run, inlier_mask, poses, n_inliers, status, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// The whole chain as ONE call on device-resident buffers, replayed from a CUDA graph: decode -> samples -> RANSAC state ->
// [minimal solver (4 kernels) -> scoring -> adaptive-stop replay] per wave -> final solve is 10+ dependent launches whose
// shapes depend only on the arguments, so the first call with a given argument set runs eagerly (sizing the workspaces,
// setting the function attributes) and captures the same enqueue into a graph; later calls are one cudaGraphLaunch.
// ---------------------------------------------------------------------------------------------------------------
static int gws_reserve(zp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->gws_bytes) return 0;
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    zp_drop_graphs(ctx);                                                        // they point into the old buffer
    if (ctx->gws) cudaFree(ctx->gws);
    ctx->gws = nullptr; ctx->gws_bytes = 0;
    ZP_CUDA(ctx, cudaMalloc(&ctx->gws, bytes + bytes / 4 + 4096));
    ctx->gws_bytes = bytes + bytes / 4 + 4096;
    return 0;
}

int zp_pose_batch_device(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4], int mask_ch,
                         int bit0_ch, int n_bits, int ignore_bit, const uint8_t* ext_mask, const double* bbox, const double* K,
                         const int32_t* obj_ids, int obj_default, int H, int m, float thr_px, double confidence, int sampler,
                         uint64_t seed, int select_mode, int final_mode, double* poses, int32_t* n_inliers, int32_t* status,
                         double* records, int use_graph, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!logits || !strides || !bbox || !K || !poses || !n_inliers || !status) ZP_FAIL(ctx, -1, "zp_pose_batch_device: null argument");
    if (S <= 0 || S > 4096) ZP_FAIL(ctx, -1, "zp_pose_batch_device: bad S %d", S);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int cap = ((S * S + 3) / 4) * 4;
    const size_t b_corr = align256((size_t)B * 5 * cap * 4), b_cnt = align256((size_t)B * 4);
    if (gws_reserve(ctx, b_corr + b_cnt)) return -2;
    float* d_corr = (float*)ctx->gws;
    int32_t* d_cnt = (int32_t*)((char*)ctx->gws + b_corr);
    cudaStream_t st = (cudaStream_t)stream;
    auto enqueue = [&]() -> int {
        if (int r = zp_decode(ctx, logits, dtype, B, S, strides, mask_ch, bit0_ch, n_bits, ignore_bit, ext_mask, bbox, obj_ids,
                              obj_default, nullptr, d_corr, cap, d_cnt, st)) return r;
        return ransac_impl(ctx, d_corr, cap, d_cnt, K, nullptr, B, H, m, thr_px, confidence, sampler, seed, select_mode, final_mode,
                           nullptr, nullptr, nullptr, nullptr, nullptr, poses, n_inliers, status, records, st);
    };
    // (the legacy default stream cannot be captured: calls on it stay eager)
    if (!use_graph || ctx->timing || st == nullptr) return enqueue();
    ZpGraphKey key;
    memset(&key, 0, sizeof key);
    key.p[0] = logits; key.p[1] = ext_mask; key.p[2] = bbox; key.p[3] = K; key.p[4] = obj_ids; key.p[5] = poses; key.p[6] = n_inliers;
    key.p[7] = status; key.p[8] = records; key.p[9] = stream;
    for (int q = 0; q < 4; q++) key.s[q] = strides[q];
    const int iv[16] = {dtype, B, S, mask_ch, bit0_ch, n_bits, ignore_bit, obj_default, H, m, sampler, select_mode, final_mode,
                        ctx->solver, ctx->n_waves, ctx->rs_no_recount * 10000000 + ctx->fin_force * 100000 + ctx->force_decode_path * 1000 + ctx->decode_rpc};
    for (int q = 0; q < 16; q++) key.i[q] = iv[q];
    for (int q = 0; q < 16; q++) key.w[q] = q < ctx->n_waves ? ctx->wave_sizes[q] : 0;
    key.f = thr_px; key.c = confidence; key.seed = seed;
    for (auto& g : ctx->graphs)
        if (!memcmp(&g.key, &key, sizeof key)) {
            ZP_CUDA(ctx, cudaGraphLaunch(g.exec, st));
            ctx->launches += g.kernels;
            return 0;
        }
    // first use: an eager run (this call's result; grows workspaces, sets attributes), then the capture of the same enqueue
    if (int r = enqueue()) return r;
    if (ctx->graphs.size() >= 32) return 0;                         // cache full: stay eager for new argument sets
    const int64_t before = ctx->launches;
    cudaGraph_t graph = nullptr;
    ZP_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue();
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    const int kernels = (int)(ctx->launches - before);
    ctx->launches = before;                                         // nothing ran during the capture
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc) return rc;
        ZP_FAIL(ctx, -2, "zp_pose_batch_device: graph capture failed: %s", cudaGetErrorString(ce));
    }
    ZpGraph g;
    g.key = key; g.kernels = kernels; g.exec = nullptr;
    ce = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) ZP_FAIL(ctx, -2, "zp_pose_batch_device: cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    ctx->graphs.push_back(g);
    return 0;
}

int zp_pose_batch_host_async(zp_ctx* ctx, const void* h_logits, int dtype, int B, int C, int S, int mask_ch, int bit0_ch,
                             int n_bits, int ignore_bit, const double* h_bbox, const double* h_K, const int32_t* h_obj_ids,
                             int obj_default, int H, int m, float thr_px, double confidence, int sampler, uint64_t seed,
                             int select_mode, int final_mode, double* h_poses, int32_t* h_n_inliers, int32_t* h_status) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!h_logits || !h_bbox || !h_K || !h_poses || !h_n_inliers || !h_status) ZP_FAIL(ctx, -1, "zp_pose_batch_host: null argument");
    if (dtype != ZP_DTYPE_F32 && dtype != ZP_DTYPE_BF16) ZP_FAIL(ctx, -1, "zp_pose_batch_host: bad dtype");
    if (B < 0 || C <= 0 || S <= 0) ZP_FAIL(ctx, -1, "zp_pose_batch_host: bad B/C/S %d/%d/%d", B, C, S);
    if (n_bits < 1 || n_bits > 16 || ignore_bit < 0 || ignore_bit >= n_bits) ZP_FAIL(ctx, -1, "zp_pose_batch_host: bad n_bits/ignore_bit");
    // the staged device copy holds exactly C channels: a layout that reaches past them would read out of bounds
    if (mask_ch < 0 || mask_ch >= C || bit0_ch < 0 || bit0_ch + (n_bits - ignore_bit) > C)
        ZP_FAIL(ctx, -1, "zp_pose_batch_host: mask_ch %d / bit0_ch %d + %d bit planes do not fit the %d channels of the logits",
                mask_ch, bit0_ch, n_bits - ignore_bit, C);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    ZpRange range("zp_pose_batch_host: H2D + decode + RANSAC + D2H");
    const size_t esz = dtype == ZP_DTYPE_F32 ? 4 : 2;
    const int cap = ((S * S + 3) / 4) * 4;
    const size_t b_log = align256((size_t)B * C * S * S * esz), b_box = align256((size_t)B * 4 * 8), b_K = align256((size_t)B * 9 * 8),
                 b_obj = align256((size_t)B * 4), b_corr = align256((size_t)B * 5 * cap * 4), b_cnt = align256((size_t)B * 4),
                 b_pose = align256((size_t)B * 12 * 8), b_ni = align256((size_t)B * 4), b_st = align256((size_t)B * 4);
    if (hws_reserve(ctx, b_log + b_box + b_K + b_obj + b_corr + b_cnt + b_pose + b_ni + b_st)) return -2;
    char* p = (char*)ctx->hws;
    void* d_log = p; p += b_log;
    double* d_box = (double*)p; p += b_box;
    double* d_K = (double*)p; p += b_K;
    int32_t* d_obj = (int32_t*)p; p += b_obj;
    float* d_corr = (float*)p; p += b_corr;
    int32_t* d_cnt = (int32_t*)p; p += b_cnt;
    double* d_pose = (double*)p; p += b_pose;
    int32_t* d_ni = (int32_t*)p; p += b_ni;
    int32_t* d_st = (int32_t*)p;
    cudaStream_t st = ctx->own_stream;
    ZP_CUDA(ctx, cudaMemcpyAsync(d_log, h_logits, (size_t)B * C * S * S * esz, cudaMemcpyHostToDevice, st));
    ZP_CUDA(ctx, cudaMemcpyAsync(d_box, h_bbox, (size_t)B * 4 * 8, cudaMemcpyHostToDevice, st));
    ZP_CUDA(ctx, cudaMemcpyAsync(d_K, h_K, (size_t)B * 9 * 8, cudaMemcpyHostToDevice, st));
    if (h_obj_ids) ZP_CUDA(ctx, cudaMemcpyAsync(d_obj, h_obj_ids, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    int64_t strides[4] = {(int64_t)C * S * S, (int64_t)S * S, S, 1};
    if (int r = zp_decode(ctx, d_log, dtype, B, S, strides, mask_ch, bit0_ch, n_bits, ignore_bit, nullptr, d_box,
                          h_obj_ids ? d_obj : nullptr, obj_default, nullptr, d_corr, cap, d_cnt, st)) return r;
    if (int r = zp_ransac(ctx, d_corr, cap, d_cnt, d_K, nullptr, B, H, m, thr_px, confidence, sampler, seed, select_mode,
                          final_mode, nullptr, nullptr, nullptr, nullptr, nullptr, d_pose, d_ni, d_st, st)) return r;
    ZP_CUDA(ctx, cudaMemcpyAsync(h_poses, d_pose, (size_t)B * 12 * 8, cudaMemcpyDeviceToHost, st));
    ZP_CUDA(ctx, cudaMemcpyAsync(h_n_inliers, d_ni, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    ZP_CUDA(ctx, cudaMemcpyAsync(h_status, d_st, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    return 0;
}

int zp_sync(zp_ctx* ctx) {
    if (!ctx) return -1;
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    ZP_CUDA(ctx, cudaStreamSynchronize(ctx->own_stream));
    return 0;
}

int zp_pose_batch_host(zp_ctx* ctx, const void* h_logits, int dtype, int B, int C, int S, int mask_ch, int bit0_ch,
                       int n_bits, int ignore_bit, const double* h_bbox, const double* h_K, const int32_t* h_obj_ids,
                       int obj_default, int H, int m, float thr_px, double confidence, int sampler, uint64_t seed,
                       int select_mode, int final_mode, double* h_poses, int32_t* h_n_inliers, int32_t* h_status) {
    if (int r = zp_pose_batch_host_async(ctx, h_logits, dtype, B, C, S, mask_ch, bit0_ch, n_bits, ignore_bit, h_bbox, h_K,
                                         h_obj_ids, obj_default, H, m, thr_px, confidence, sampler, seed, select_mode,
                                         final_mode, h_poses, h_n_inliers, h_status)) return r;
    if (B == 0) return 0;
    return zp_sync(ctx);
}

int zp_debug_clocks(zp_ctx* ctx, int64_t* out16) {   // 24 slots
    if (!ctx || !out16) return -1;
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    return zp_read_debug_clocks((long long*)out16);
}

int zp_fp32_peak_probe(zp_ctx* ctx, int iters, double* out_tflops) {
    if (!ctx || !out_tflops) return -1;
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    return zp_launch_fma_probe(ctx, iters, 0, out_tflops);
}

int zp_fp64_peak_probe(zp_ctx* ctx, int iters, double* out_tflops) {
    if (!ctx || !out_tflops) return -1;
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    return zp_launch_dfma_probe(ctx, iters, out_tflops);
}

int zp_fp32x2_peak_probe(zp_ctx* ctx, int iters, double* out_tflops) {
    if (!ctx || !out_tflops) return -1;
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    return zp_launch_fma_probe(ctx, iters, 1, out_tflops);
}

}  // extern "C"
