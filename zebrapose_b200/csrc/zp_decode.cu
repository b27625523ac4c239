// Kernel 1: per-pixel decode + stable stream compaction (HBM-bandwidth bound).
//
// Replaces, for a whole batch of crops and without a host copy (reference file:line, /root/reference/zebrapose):
//   common_ops.py:5-19                       sigmoid(x) > 0.5         -> float32(x) > 0
//   class_id_encoder_decoder.py:17-28        bits -> class id         -> MSB-first pack of the first nb planes
//   CNN_output_to_pose.py:111,53-64          mask.nonzero() + dict    -> row-major stable compaction + table gather
//   CNN_output_to_pose.py:34-50              pixel -> original image  -> float64 w/S*x + x0, truncation
//
// Layout / mapping (fast path): one thread-block CLUSTER per crop.  Each CTA (512 threads) owns a contiguous
// run of 512*PPT pixels (PPT = 16 B / sizeof(logit): 4 for fp32, 8 for bf16); a thread reads PPT adjacent pixels
// of every plane with one 128-bit streaming load (each plane row is contiguous -> fully coalesced), packs the code
// in registers, and the masked-pixel ranks come from warp shuffles + one shared-memory scan; the CTA totals are
// exchanged through distributed shared memory so the output order is the crop's row-major order with no atomics
// and no second pass over HBM.  The point table (float4[2^nb], <= 1 MB) is gathered from L2.
//
// Algorithmic HBM bytes per crop: (1 + nb) * S*S * sizeof(logit) read + 20 * M + 4 written.
#include <cooperative_groups.h>
#include "zp_common.cuh"

namespace cg = cooperative_groups;

constexpr int DEC_THREADS = 512;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int DEC_MAX_CLUSTER = 8;

struct DecodeArgs {
    const void* logits;
    int64_t sb, sc, sh, sw;      // strides in elements
    int B, S, mask_ch, bit0_ch, nb;
    const uint8_t* ext_mask;
    const double* bbox;
    const int32_t* obj_ids;
    int obj_default;
    const float4* const* tables;
    uint16_t* codes;
    float* corr;
    int cap;
    int32_t* counts;
    int32_t* chunk_counts;       // generic path only: [B, n_chunks]
    int n_chunks;
};

__device__ __forceinline__ bool pos_f32(uint32_t bits) { return __uint_as_float(bits) > 0.0f; }
// bf16 > 0  <=>  sign clear, not zero, not NaN  <=>  bits in [0x0001, 0x7F80]
__device__ __forceinline__ bool pos_bf16(uint32_t h) { return (uint32_t)((h & 0xFFFFu) - 1u) < 0x7F80u; }

template <int DT> struct Px;   // pixels per 16-byte load
template <> struct Px<ZP_DTYPE_F32> { static constexpr int N = 4; static constexpr int ESZ = 4; };
template <> struct Px<ZP_DTYPE_BF16> { static constexpr int N = 8; static constexpr int ESZ = 2; };

// bit i of the result = (pixel i of the vector > 0)
template <int DT> __device__ __forceinline__ uint32_t positive_bits(const uint4& v) {
    if (DT == ZP_DTYPE_F32) {
        return (uint32_t)pos_f32(v.x) | ((uint32_t)pos_f32(v.y) << 1) | ((uint32_t)pos_f32(v.z) << 2) |
               ((uint32_t)pos_f32(v.w) << 3);
    } else {
        return (uint32_t)pos_bf16(v.x) | ((uint32_t)pos_bf16(v.x >> 16) << 1) | ((uint32_t)pos_bf16(v.y) << 2) |
               ((uint32_t)pos_bf16(v.y >> 16) << 3) | ((uint32_t)pos_bf16(v.z) << 4) |
               ((uint32_t)pos_bf16(v.z >> 16) << 5) | ((uint32_t)pos_bf16(v.w) << 6) |
               ((uint32_t)pos_bf16(v.w >> 16) << 7);
    }
}

// CNN_output_to_pose.py:41-48 in float64: ratio = w / S; int(ratio * p + x0) (truncate toward zero).
// __dmul_rn/__dadd_rn keep the two roundings of numpy (no FMA contraction).
__device__ __forceinline__ float remap_coord(double extent, double origin, int S, int p) {
    double ratio = extent / (double)S;
    double v = __dadd_rn(__dmul_rn(ratio, (double)p), origin);
    return (float)__double2ll_rz(v);
}

// -------------------------------------------------------------------------------------------------------------
// Fast path: cluster per crop, vector loads.  grid.x = B * cluster_size.
// -------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(DEC_THREADS, 2) zp_decode_cluster_kernel(DecodeArgs a) {
    constexpr int PPT = Px<DT>::N;
    constexpr int ESZ = Px<DT>::ESZ;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.x / csize;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.S, N = S * S;

    __shared__ float s_x[1024], s_y[1024];     // remapped coordinates per column / row (S <= 1024 on this path)
    __shared__ int s_warp[DEC_WARPS];
    __shared__ int s_total;                    // read by the other CTAs of the cluster (DSMEM)

    {   // per-crop coordinate LUTs
        const double* bb = a.bbox + 4 * (size_t)b;
        double x0 = bb[0], y0 = bb[1], w = bb[2], h = bb[3];
        for (int i = tid; i < S; i += DEC_THREADS) {
            s_x[i] = remap_coord(w, x0, S, i);
            s_y[i] = remap_coord(h, y0, S, i);
        }
    }

    const int p0 = ((int)rank * DEC_THREADS + tid) * PPT;       // first pixel of this thread
    const bool active = p0 < N;
    const int row = active ? p0 / S : 0, col = active ? p0 - row * S : 0;
    const char* base = (const char*)a.logits + ((size_t)b * a.sb + (size_t)row * a.sh + col) * ESZ;
    const size_t plane = (size_t)a.sc * ESZ;

    // ---- issue the mask load and the first half of the bit planes together (memory-level parallelism)
    uint32_t mbits = 0;
    uint4 v[8];
    const int nb = a.nb;
    if (active) {
        uint4 mv = make_uint4(0, 0, 0, 0);
        if (a.ext_mask == nullptr) mv = zp_ldg_stream(base + (size_t)a.mask_ch * plane);
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < nb) v[i] = zp_ldg_stream(base + (size_t)(a.bit0_ch + i) * plane);
        if (a.ext_mask == nullptr) {
            mbits = positive_bits<DT>(mv);
        } else {
            const uint8_t* em = a.ext_mask + (size_t)b * N + p0;
#pragma unroll
            for (int j = 0; j < PPT; j++) mbits |= (uint32_t)(em[j] != 0) << j;
        }
    }
    // ---- ranks of the masked pixels: thread -> warp -> CTA -> cluster
    const int cnt = __popc(mbits);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < DEC_WARPS ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < DEC_WARPS; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < DEC_WARPS) s_warp[lane] = wi - w;      // exclusive warp offsets
        if (lane == DEC_WARPS - 1) s_total = wi;
    }
    __syncthreads();
    if (csize > 1) cluster.barrier_arrive();              // release s_total; the wait is after the code planes

    // ---- pack the code while the exchange is in flight
    uint32_t code[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j++) code[j] = 0;
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < nb) {
                uint32_t pb = positive_bits<DT>(v[i]);
#pragma unroll
                for (int j = 0; j < PPT; j++) code[j] |= ((pb >> j) & 1u) << (nb - 1 - i);
            }
        if (nb > 8) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (8 + i < nb) v[i] = zp_ldg_stream(base + (size_t)(a.bit0_ch + 8 + i) * plane);
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (8 + i < nb) {
                    uint32_t pb = positive_bits<DT>(v[i]);
#pragma unroll
                    for (int j = 0; j < PPT; j++) code[j] |= ((pb >> j) & 1u) << (nb - 9 - i);
                }
        }
        if (a.codes) {
            uint16_t* cp = a.codes + (size_t)b * N + p0;
            if (PPT == 4) {
                *(uint2*)cp = make_uint2(code[0] | (code[1] << 16), code[2] | (code[3] << 16));
            } else {
                *(uint4*)cp = make_uint4(code[0] | (code[1] << 16), code[2] | (code[3] << 16),
                                         code[4 % PPT] | (code[5 % PPT] << 16), code[6 % PPT] | (code[7 % PPT] << 16));
            }
        }
    }

    // ---- cluster-wide exclusive prefix of the CTA totals
    int cta_base = 0;
    if (csize > 1) {
        cluster.barrier_wait();
        for (unsigned r = 0; r < rank; r++) cta_base += *cluster.map_shared_rank(&s_total, r);
        cluster.barrier_arrive();                         // peers may retire once everybody has read
    }
    if (rank == csize - 1 && tid == 0) a.counts[b] = cta_base + s_total;

    // ---- gather the 3D points and emit this thread's correspondences in row-major order
    if (cnt) {
        const int obj = a.obj_ids ? a.obj_ids[b] : a.obj_default;
        const float4* tab = a.tables[obj];
        int pos = cta_base + s_warp[warp] + incl - cnt;
        float* cb = a.corr + (size_t)b * 5 * a.cap;
        const float yv = s_y[row];
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            if ((mbits >> j) & 1u) {
                if (pos < a.cap) {
                    float4 P = __ldg(tab + code[j]);
                    cb[pos] = s_x[col + j];
                    cb[a.cap + pos] = yv;
                    cb[2 * (size_t)a.cap + pos] = P.x;
                    cb[3 * (size_t)a.cap + pos] = P.y;
                    cb[4 * (size_t)a.cap + pos] = P.z;
                }
                pos++;
            }
        }
    }
    if (csize > 1) cluster.barrier_wait();
}

// -------------------------------------------------------------------------------------------------------------
// Generic path (any strides / crop size): scalar loads, chunked two-kernel compaction.
// chunk = DEC_THREADS consecutive pixels; kernel A counts masked pixels per chunk, kernel B emits.
// -------------------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ bool load_positive(const DecodeArgs& a, int b, int ch, int row, int col) {
    size_t off = (size_t)b * a.sb + (size_t)ch * a.sc + (size_t)row * a.sh + (size_t)col * a.sw;
    if (DT == ZP_DTYPE_F32) return ((const float*)a.logits)[off] > 0.0f;
    return pos_bf16(((const uint16_t*)a.logits)[off]);
}

template <int DT>
__global__ void __launch_bounds__(DEC_THREADS) zp_decode_generic_kernel(DecodeArgs a, int phase) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.S, N = S * S;
    const int p = chunk * DEC_THREADS + tid;
    const bool active = p < N;
    const int row = active ? p / S : 0, col = active ? p - row * S : 0;
    __shared__ int s_warp[DEC_WARPS];
    __shared__ int s_base;
    bool m = false;
    if (active) m = a.ext_mask ? a.ext_mask[(size_t)b * N + p] != 0 : load_positive<DT>(a, b, a.mask_ch, row, col);
    unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0) s_warp[warp] = __popc(bal);
    if (phase == 1 && tid == 0) {
        int base = 0;
        for (int c = 0; c < chunk; c++) base += a.chunk_counts[(size_t)b * a.n_chunks + c];
        s_base = base;
    }
    __syncthreads();
    if (phase == 0) {
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < DEC_WARPS; w++) t += s_warp[w];
            a.chunk_counts[(size_t)b * a.n_chunks + chunk] = t;
        }
        return;
    }
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += s_warp[w];
    if (chunk == a.n_chunks - 1 && tid == DEC_THREADS - 1) {
        int t = 0;
        for (int w = 0; w < DEC_WARPS; w++) t += s_warp[w];
        a.counts[b] = s_base + t;
    }
    if (!active) return;
    uint32_t code = 0;
    for (int i = 0; i < a.nb; i++)
        code |= (uint32_t)load_positive<DT>(a, b, a.bit0_ch + i, row, col) << (a.nb - 1 - i);
    if (a.codes) a.codes[(size_t)b * N + p] = (uint16_t)code;
    if (m) {
        int pos = s_base + woff + __popc(bal & ((1u << lane) - 1u));
        if (pos < a.cap) {
            const double* bb = a.bbox + 4 * (size_t)b;
            const int obj = a.obj_ids ? a.obj_ids[b] : a.obj_default;
            float4 P = __ldg(a.tables[obj] + code);
            float* cb = a.corr + (size_t)b * 5 * a.cap;
            cb[pos] = remap_coord(bb[2], bb[0], S, col);
            cb[a.cap + pos] = remap_coord(bb[3], bb[1], S, row);
            cb[2 * (size_t)a.cap + pos] = P.x;
            cb[3 * (size_t)a.cap + pos] = P.y;
            cb[4 * (size_t)a.cap + pos] = P.z;
        }
    }
}

template <int DT>
static int launch_cluster(zp_ctx* ctx, const DecodeArgs& a, int csize, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.B * csize));
    cfg.blockDim = dim3(DEC_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ZP_CUDA(ctx, cudaLaunchKernelEx(&cfg, zp_decode_cluster_kernel<DT>, a));
    ctx->launches++;
    return 0;
}

int zp_launch_decode(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4],
                     int mask_ch, int bit0_ch, int nb, const uint8_t* ext_mask, const double* bbox,
                     const int32_t* obj_ids, int obj_default, uint16_t* codes, float* corr, int cap,
                     int32_t* counts, cudaStream_t st) {
    DecodeArgs a;
    a.logits = logits; a.sb = strides[0]; a.sc = strides[1]; a.sh = strides[2]; a.sw = strides[3];
    a.B = B; a.S = S; a.mask_ch = mask_ch; a.bit0_ch = bit0_ch; a.nb = nb;
    a.ext_mask = ext_mask; a.bbox = bbox; a.obj_ids = obj_ids; a.obj_default = obj_default;
    a.tables = (const float4* const*)ctx->d_table_ptrs;
    a.codes = codes; a.corr = corr; a.cap = cap; a.counts = counts;
    a.chunk_counts = nullptr; a.n_chunks = 0;
    const int ppt = dtype == ZP_DTYPE_F32 ? 4 : 8;
    const int esz = dtype == ZP_DTYPE_F32 ? 4 : 2;
    const int N = S * S;
    const int per_cta = DEC_THREADS * ppt;
    const int ctas = (N + per_cta - 1) / per_cta;
    bool vec_ok = a.sw == 1 && S % ppt == 0 && S <= 1024 && a.sh % ppt == 0 && a.sc % ppt == 0 && a.sb % ppt == 0 &&
                  ((uintptr_t)logits % 16) == 0 && ctas <= DEC_MAX_CLUSTER &&
                  (codes == nullptr || ((uintptr_t)codes % 16) == 0);
    (void)esz;
    if (vec_ok) {
        int csize = 1;
        while (csize < ctas) csize <<= 1;
        if (dtype == ZP_DTYPE_F32) return launch_cluster<ZP_DTYPE_F32>(ctx, a, csize, st);
        return launch_cluster<ZP_DTYPE_BF16>(ctx, a, csize, st);
    }
    // generic path
    a.n_chunks = (N + DEC_THREADS - 1) / DEC_THREADS;
    size_t need = (size_t)B * a.n_chunks * sizeof(int32_t);
    if (zp_ws_reserve(ctx, need)) return -2;
    a.chunk_counts = (int32_t*)ctx->ws;
    dim3 grid(a.n_chunks, B);
    for (int phase = 0; phase < 2; phase++) {
        if (dtype == ZP_DTYPE_F32) zp_decode_generic_kernel<ZP_DTYPE_F32><<<grid, DEC_THREADS, 0, st>>>(a, phase);
        else zp_decode_generic_kernel<ZP_DTYPE_BF16><<<grid, DEC_THREADS, 0, st>>>(a, phase);
        ZP_CHECK_LAUNCH(ctx, "zp_decode_generic_kernel");
    }
    return 0;
}

// -------------------------------------------------------------------------------------------------------------
// stand-alone helpers (reference signatures that callers may still use outside the batched path)
// -------------------------------------------------------------------------------------------------------------
__global__ void zp_remap_pixels_kernel(const int64_t* __restrict__ px, int64_t N, double x0, double y0, double w,
                                       double h, int S, int64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double rx = w / (double)S, ry = h / (double)S;
    out[2 * i] = __double2ll_rz(__dadd_rn(__dmul_rn(rx, (double)px[2 * i]), x0));
    out[2 * i + 1] = __double2ll_rz(__dadd_rn(__dmul_rn(ry, (double)px[2 * i + 1]), y0));
}

__global__ void zp_codes_to_ids_kernel(const double* __restrict__ bits, int64_t N, int L, int base,
                                       double* __restrict__ ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double acc = 0;
    for (int c = 0; c < L; c++) {            // same left-to-right float64 accumulation as the reference loop
        double wgt = 1;
        for (int q = 0; q < L - 1 - c; q++) wgt *= base;
        acc = __dadd_rn(acc, __dmul_rn(bits[i * L + c], wgt));
    }
    ids[i] = acc;
}

int zp_launch_remap_pixels(zp_ctx* ctx, const int64_t* px, int64_t N, const double* bb, int S, int64_t* out, cudaStream_t st) {
    if (N == 0) return 0;
    zp_remap_pixels_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(px, N, bb[0], bb[1], bb[2], bb[3], S, out);
    ZP_CHECK_LAUNCH(ctx, "zp_remap_pixels_kernel");
    return 0;
}

int zp_launch_codes_to_ids(zp_ctx* ctx, const double* bits, int64_t N, int L, int base, double* ids, cudaStream_t st) {
    if (N == 0) return 0;
    zp_codes_to_ids_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(bits, N, L, base, ids);
    ZP_CHECK_LAUNCH(ctx, "zp_codes_to_ids_kernel");
    return 0;
}
