// Fused network tail: the reference ends its network with
//     x = self.conv_1x1_4(torch.cat([x, x_128], 1))            (zebrapose/model/aspp.py:58,112; 256+64 -> 17 channels)
// writes the [B,17,128,128] logits to HBM, and the pose path reads them back (common_ops.py:5-19).  Here the 1x1
// convolution, the sigmoid > 0.5 threshold and the MSB-first bit packing are ONE kernel: the two feature tensors are read
// once (no concatenation copy), the logits never exist in memory, and the epilogue emits the decode path's compact
// intermediate (2-byte code + 1 mask bit per pixel) which zp_decode_emit_kernel turns into correspondence lists.
//
// A 1x1 convolution over channels-last activations is a dense contraction [pixels x C_in] . [C_in x 17]: this is the one
// GEMM-shaped step of the path, so it runs on the 5th-generation tensor cores:
//   * warp 0     TMA producer: 128-pixel x 64-channel bf16 boxes (128-byte swizzle) into a 6-stage shared-memory ring,
//                from x for the first C1/64 k-blocks and from the skip tensor for the rest -- the "cat" is just which
//                tensor map a k-block uses; the padded weight matrix [32 x C_in] is loaded once per CTA;
//   * warp 1     allocates TMEM and issues tcgen05.mma (cta_group::1, kind::f16, M = 128 pixels, N = 32, K = 16) from one
//                lane; tcgen05.commit releases ring slots and publishes the accumulator;
//   * warps 2-5  epilogue: tcgen05.ld of the pixel's 32 fp32 accumulators, + bias, x > 0, bit-reverse pack, coalesced
//                2-byte code stores and the mask ballot bytes; two TMEM accumulator stages overlap it with the next tile.
// Persistent: one CTA per SM walks pixel tiles round-robin.  HBM-bound by design: 2*C_in bytes per pixel in (10.5 MB per
// 128x128 crop at C_in = 320), 2.125 bytes per pixel out.
#include <cuda.h>
#include <algorithm>
#include <cmath>
#include "zp_common.cuh"

int zp_launch_emit_codes(zp_ctx* ctx, int B, int S, const double* bbox, const int32_t* obj_ids, int obj_default,
                         const uint16_t* codes, const uint32_t* maskw, float* corr, int cap, int32_t* counts, cudaStream_t st);

constexpr int HD_TILE_M = 128;                 // pixels per tile = UMMA M
constexpr int HD_N = 32;                       // output channels padded to the UMMA N granule (17 used)
constexpr int HD_ROW_BYTES = 128;               // one k-block = one 128-byte swizzle row per pixel: 64 bf16 or 32 fp32 channels
constexpr int HD_KSTEPS = 4;                    // tcgen05.mma per k-block: 4 x (K = 16 bf16 | K = 8 tf32) = 4 x 32 bytes
constexpr int HD_STAGES = 6;
constexpr int HD_A_BYTES = HD_TILE_M * HD_ROW_BYTES;   // 16 KB
constexpr int HD_W_BYTES = HD_N * HD_ROW_BYTES;        // 4 KB per k-block
constexpr int HD_MAX_KB = 16;                          // C_in <= 1024 (bf16) | 512 (fp32)
constexpr int HD_THREADS = 192;
constexpr int HD_TMEM_COLS = 64;                       // 2 accumulator stages x 32 columns

struct HeadParams {
    int n_tiles, kb_x, kb_total, kb_ch;      // kb_ch = channels per k-block (64 bf16 | 32 fp32)
    uint32_t idesc;
    int mask_ch, bit0_ch, nb;
    uint16_t* codes;                 // [n_tiles * 128]
    uint8_t* maskb;                  // mask ballot words as bytes: [n_tiles][4 words][4 bytes]
    float bias[HD_N];
};

__device__ __forceinline__ uint32_t hd_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hd_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hd_smem(bar)), "r"(count));
}
__device__ __forceinline__ void hd_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hd_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hd_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hd_smem(bar)) : "memory");
}
__device__ __forceinline__ void hd_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "HD_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra HD_DONE;\n\t"
        "bra HD_WAIT;\n\t"
        "HD_DONE:\n\t}" ::"r"(hd_smem(bar)), "r"(parity) : "memory");
}
// 2-D TMA tile load (SASS: UTMALDG), completion on an mbarrier
__device__ __forceinline__ void hd_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(hd_smem(dst)), "l"(map), "r"(hd_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO = 64 x 16 B),
// LBO = 1 (ignored for swizzled K-major), version 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t hd_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (kind::f16): D = F32 [4,6) = 1, A = BF16 [7,10) = 1, B = BF16 [10,13) = 1, both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t hd_idesc(uint32_t fmt) {     // fmt: 1 = BF16, 2 = TF32 (fp32 operands, 10-bit mantissa products, fp32 accumulate)
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(HD_N >> 3) << 17) | ((uint32_t)(HD_TILE_M >> 4) << 24);
}

template <bool TF32>
__device__ __forceinline__ void hd_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if (TF32)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void hd_commit(uint64_t* bar) {      // arrives on `bar` when all MMAs issued so far have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hd_smem(bar)) : "memory");
}

template <bool TF32>
__global__ void __launch_bounds__(HD_THREADS, 1)
zp_head_codes_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_s,
                     const __grid_constant__ CUtensorMap map_w, const __grid_constant__ HeadParams p) {
    extern __shared__ __align__(1024) uint8_t hd_smem_raw[];
    // carve: ring of A tiles | weights | barriers   (the dynamic base is rounded up to 1024 B: swizzle atoms need it)
    uint8_t* base = (uint8_t*)(((uintptr_t)hd_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = base;
    uint8_t* s_w = base + HD_STAGES * HD_A_BYTES;
    uint64_t* bars = (uint64_t*)(s_w + HD_MAX_KB * HD_W_BYTES);
    uint64_t* full = bars;                       // [HD_STAGES]  TMA -> MMA
    uint64_t* empty = bars + HD_STAGES;          // [HD_STAGES]  MMA -> TMA
    uint64_t* tfull = bars + 2 * HD_STAGES;      // [2]          MMA -> epilogue
    uint64_t* tempty = tfull + 2;                // [2]          epilogue -> MMA
    uint64_t* wfull = tempty + 2;                // [1]
    uint32_t* s_tmem = (uint32_t*)(wfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_s) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int i = 0; i < HD_STAGES; i++) { hd_mbar_init(&full[i], 1); hd_mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { hd_mbar_init(&tfull[i], 1); hd_mbar_init(&tempty[i], 4); }
        hd_mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {          // TMEM: 64 columns x 128 lanes of fp32 (two accumulator stages)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hd_smem(s_tmem)), "n"(HD_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            hd_mbar_expect_tx(wfull, (uint32_t)p.kb_total * HD_W_BYTES);
            for (int kb = 0; kb < p.kb_total; kb++) hd_tma_2d(s_w + kb * HD_W_BYTES, &map_w, kb * p.kb_ch, 0, wfull);
        }
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < p.kb_total; kb++) {
                hd_mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    hd_mbar_expect_tx(&full[stage], HD_A_BYTES);
                    if (kb < p.kb_x) hd_tma_2d(s_a + stage * HD_A_BYTES, &map_x, kb * p.kb_ch, tile * HD_TILE_M, &full[stage]);
                    else hd_tma_2d(s_a + stage * HD_A_BYTES, &map_s, (kb - p.kb_x) * p.kb_ch, tile * HD_TILE_M, &full[stage]);
                }
                __syncwarp();
                if (++stage == HD_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one lane) =====
        hd_mbar_wait(wfull, 0);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            hd_mbar_wait(&tempty[acc], acc_phase ^ 1);              // epilogue has drained this accumulator stage
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < p.kb_total; kb++) {
                hd_mbar_wait(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint64_t da = hd_desc(hd_smem(s_a + stage * HD_A_BYTES));
                    const uint64_t db = hd_desc(hd_smem(s_w + kb * HD_W_BYTES));
#pragma unroll
                    for (int k = 0; k < HD_KSTEPS; k++)               // +32 bytes (2 x 16 B) per K step inside the swizzle row
                        hd_mma<TF32>(tmem_base + acc * HD_N, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0);
                    hd_commit(&empty[stage]);                         // slot free once these MMAs have read it
                    if (kb == p.kb_total - 1) hd_commit(&tfull[acc]); // accumulator complete
                }
                __syncwarp();
                if (++stage == HD_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32 (w % 4), +32) =====
        const int quarter = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        const uint32_t field_mask = p.nb >= 32 ? 0xffffffffu : ((1u << p.nb) - 1u);
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            hd_mbar_wait(&tfull[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * HD_N);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) hd_mbar_arrive(&tempty[acc]);
            // logit_c = acc_c + bias_c; bit = logit > 0 (common_ops.py:5-19: sigmoid(x) > 0.5; NaN -> 0)
            uint32_t pos = 0;
#pragma unroll
            for (int c = 0; c < HD_N; c++) pos |= (uint32_t)((__uint_as_float(v[c]) + p.bias[c]) > 0.0f) << c;
            // channel bit0_ch is the MOST significant code bit (class_id_encoder_decoder.py:26)
            const uint32_t field = (pos >> p.bit0_ch) & field_mask;
            const uint32_t code = __brev(field) >> (32 - p.nb);
            p.codes[(size_t)tile * HD_TILE_M + quarter * 32 + lane] = (uint16_t)code;
            // mask ballot: pixel 4i + j of the 128-pixel segment -> bit i of word j; this warp owns byte `quarter` of each word
            const uint32_t bal = __ballot_sync(0xffffffffu, (pos >> p.mask_ch) & 1u);
            if (lane < 4) {
                uint32_t byte = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) byte |= ((bal >> (4 * k + lane)) & 1u) << k;
                p.maskb[((size_t)tile * 4 + lane) * 4 + quarter] = (uint8_t)byte;
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(HD_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn hd_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [rows, cols] bf16 | fp32 row-major (cols contiguous), box = one 128-byte row segment x box_rows rows, 128-byte swizzle
static int hd_make_map(zp_ctx* ctx, CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int esz) {
    EncodeTiledFn fn = hd_encode_fn();
    if (!fn) ZP_FAIL(ctx, -2, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * (uint64_t)esz};
    cuuint32_t box[2] = {(cuuint32_t)(HD_ROW_BYTES / esz), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ZP_FAIL(ctx, -2, "cuTensorMapEncodeTiled failed (%d) for a [%llu x %llu] tensor", (int)r,
                                   (unsigned long long)rows, (unsigned long long)cols);
    return 0;
}

static uint16_t hd_f2bf(float f) {              // round to nearest even, like torch's .to(bfloat16)
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

static size_t hd_align(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" {

int zp_upload_head(zp_ctx* ctx, const float* weight, const float* bias, int n_out, int c_in) {
    if (!ctx) return -1;
    if (!weight || n_out < 1 || n_out > HD_N) ZP_FAIL(ctx, -1, "zp_upload_head: n_out must be 1..%d, got %d", HD_N, n_out);
    if (c_in < 64 || c_in % 64 != 0 || c_in > 512) ZP_FAIL(ctx, -1, "zp_upload_head: c_in must be a multiple of 64 up to 512, got %d", c_in);
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint16_t> w((size_t)HD_N * c_in, 0);
    for (int o = 0; o < n_out; o++)
        for (int c = 0; c < c_in; c++) w[(size_t)o * c_in + c] = hd_f2bf(weight[(size_t)o * c_in + c]);
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->head_w) cudaFree(ctx->head_w);
    ctx->head_w = nullptr;
    ZP_CUDA(ctx, cudaMalloc(&ctx->head_w, w.size() * 2));
    ZP_CUDA(ctx, cudaMemcpy(ctx->head_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
    // float32 copy for the TF32 path (fp32 activations): the tensor core reads the top 19 bits of each operand
    std::vector<float> w32((size_t)HD_N * c_in, 0.f);
    for (int o = 0; o < n_out; o++)
        for (int c = 0; c < c_in; c++) w32[(size_t)o * c_in + c] = weight[(size_t)o * c_in + c];
    if (ctx->head_w32) cudaFree(ctx->head_w32);
    ctx->head_w32 = nullptr;
    ZP_CUDA(ctx, cudaMalloc(&ctx->head_w32, w32.size() * 4));
    ZP_CUDA(ctx, cudaMemcpy(ctx->head_w32, w32.data(), w32.size() * 4, cudaMemcpyHostToDevice));
    for (int o = 0; o < HD_N; o++) ctx->head_bias[o] = (bias && o < n_out) ? bias[o] : 0.f;
    ctx->head_n_out = n_out; ctx->head_c_in = c_in;
    return 0;
}

int zp_head_decode(zp_ctx* ctx, const void* x, int c1, const void* x_skip, int c2, int dtype, int B, int S, int mask_ch, int bit0_ch,
                   int n_bits, int ignore_bit, const double* bbox, const int32_t* obj_ids, int obj_default,
                   uint16_t* codes, float* corr, int cap, int32_t* counts, void* stream) {
    if (!ctx) return -1;
    if (B == 0) return 0;
    if (!ctx->head_w) ZP_FAIL(ctx, -1, "zp_head_decode: no head weights uploaded (zp_upload_head)");
    if (!x || !bbox || !corr || !counts || B < 0) ZP_FAIL(ctx, -1, "zp_head_decode: null argument");
    if (dtype != ZP_DTYPE_F32 && dtype != ZP_DTYPE_BF16) ZP_FAIL(ctx, -1, "zp_head_decode: dtype %d not supported", dtype);
    const int esz = dtype == ZP_DTYPE_BF16 ? 2 : 4;
    const int kb_ch = HD_ROW_BYTES / esz;
    if (c2 < 0 || (c2 > 0 && !x_skip) || c1 < kb_ch || c1 % kb_ch || c2 % kb_ch || c1 + c2 != ctx->head_c_in)
        ZP_FAIL(ctx, -1, "zp_head_decode: channel split %d + %d does not match the uploaded head (c_in %d, multiples of %d)", c1, c2, ctx->head_c_in, kb_ch);
    const int nb = n_bits - ignore_bit;
    if (n_bits < 1 || n_bits > 16 || ignore_bit < 0 || nb < 1) ZP_FAIL(ctx, -1, "zp_head_decode: bad n_bits/ignore_bit");
    if (mask_ch < 0 || mask_ch >= ctx->head_n_out || bit0_ch < 0 || bit0_ch + nb > ctx->head_n_out)
        ZP_FAIL(ctx, -1, "zp_head_decode: channel layout exceeds the %d outputs of the head", ctx->head_n_out);
    const long long N = (long long)S * S;
    if (S <= 0 || S % 4 != 0 || N % HD_TILE_M != 0) ZP_FAIL(ctx, -1, "zp_head_decode: S*S must be a multiple of %d and S of 4 (S = %d)", HD_TILE_M, S);
    if (((uintptr_t)x % 16) || ((uintptr_t)x_skip % 16)) ZP_FAIL(ctx, -1, "zp_head_decode: activations must be 16-byte aligned");
    if (!obj_ids) {
        if (obj_default < 0 || obj_default >= ZP_MAX_OBJECTS || !ctx->tables[obj_default].pts)
            ZP_FAIL(ctx, -1, "zp_head_decode: no dictionary uploaded for object slot %d", obj_default);
        const ZpTable& t = ctx->tables[obj_default];
        if (t.n_bits != n_bits || t.ignore_bit != ignore_bit)
            ZP_FAIL(ctx, -1, "zp_head_decode: slot %d holds a %d-bit/ignore %d table, call asks %d/%d", obj_default, t.n_bits, t.ignore_bit, n_bits, ignore_bit);
    }
    ZP_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_px = (long long)B * N;
    const int n_tiles = (int)(n_px / HD_TILE_M);
    // workspace: codes (unless the caller wants them) | mask ballot words
    const size_t b_codes = codes ? 0 : hd_align((size_t)n_px * 2);
    const size_t b_mask = hd_align((size_t)n_tiles * 16);
    if (ctx->hdws_bytes < b_codes + b_mask) {
        ZP_CUDA(ctx, cudaDeviceSynchronize());
        if (ctx->hdws) cudaFree(ctx->hdws);
        ctx->hdws = nullptr; ctx->hdws_bytes = 0;
        ZP_CUDA(ctx, cudaMalloc(&ctx->hdws, b_codes + b_mask + 4096));
        ctx->hdws_bytes = b_codes + b_mask + 4096;
    }
    uint16_t* d_codes = codes ? codes : (uint16_t*)ctx->hdws;
    uint8_t* d_mask = (uint8_t*)ctx->hdws + b_codes;
    CUtensorMap mx, ms, mw;
    if (int r = hd_make_map(ctx, &mx, x, (uint64_t)n_px, (uint64_t)c1, HD_TILE_M, esz)) return r;
    if (c2 > 0) { if (int r = hd_make_map(ctx, &ms, x_skip, (uint64_t)n_px, (uint64_t)c2, HD_TILE_M, esz)) return r; }
    else ms = mx;
    if (int r = hd_make_map(ctx, &mw, esz == 2 ? ctx->head_w : ctx->head_w32, HD_N, (uint64_t)ctx->head_c_in, HD_N, esz)) return r;
    HeadParams p{};
    p.n_tiles = n_tiles; p.kb_x = c1 / kb_ch; p.kb_total = (c1 + c2) / kb_ch; p.kb_ch = kb_ch;
    p.idesc = hd_idesc(esz == 2 ? 1u : 2u);
    p.mask_ch = mask_ch; p.bit0_ch = bit0_ch; p.nb = nb;
    p.codes = d_codes; p.maskb = d_mask;
    for (int o = 0; o < HD_N; o++) p.bias[o] = ctx->head_bias[o];
    const int smem = HD_STAGES * HD_A_BYTES + HD_MAX_KB * HD_W_BYTES + 256 + 1024;
    static bool attr_set_dev[ZP_MAX_DEVICES] = {};          // cudaFuncSetAttribute is per device
    bool& attr_set = attr_set_dev[ctx->device % ZP_MAX_DEVICES];
    if (!attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_head_codes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_head_codes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const int grid = std::min(n_tiles, ctx->sm_count);
    ZP_TIME_BEGIN(ctx, st);
    if (esz == 2) zp_head_codes_kernel<false><<<grid, HD_THREADS, smem, st>>>(mx, ms, mw, p);
    else zp_head_codes_kernel<true><<<grid, HD_THREADS, smem, st>>>(mx, ms, mw, p);
    ZP_CHECK_LAUNCH(ctx, "zp_head_codes_kernel");
    return zp_launch_emit_codes(ctx, B, S, bbox, obj_ids, obj_default, d_codes, (const uint32_t*)d_mask, corr, cap, counts, st);
}

}  // extern "C"
