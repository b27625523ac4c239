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
constexpr int DEC_MAX_RUNS = 16;               // runs (CTAs) per crop of the cluster-free kernel: run r re-counts r mask runs

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
    unsigned long long* dbg;     // nullable profiling aid: per-CTA (start, loads done, ranks done, end) %globaltimer stamps
};
__device__ __forceinline__ unsigned long long zp_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DEC_STAMP(k) do { if (a.dbg && threadIdx.x == 0) a.dbg[4 * (size_t)blockIdx.x + (k)] = zp_gtimer(); } while (0)

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

// Emission through shared memory: every thread drops its <= PPT correspondences at their CTA-local ranks into a
// [5][DEC_THREADS * PPT] staging buffer, then the CTA copies the compacted run out with consecutive lanes on consecutive
// addresses.  Writing straight from the owning threads costs one partially filled 32-byte sector per lane and store
// (lanes are ~3 floats apart, and a sector is touched again by each of the PPT unrolled stores): ncu counted 32.7 M
// sector writes for 12.6 M correspondences, which -- not DRAM -- was what bounded the first kernels (bf16 input, half
// the bytes, took exactly as long as fp32).
template <int PPT>
__device__ __forceinline__ void emit_staged(float* s_out, uint32_t mbits, const uint32_t* code2, int local_pos,
                                            const float4* __restrict__ tab, const float* s_x, int col, float yv,
                                            int total, int cta_base, float* __restrict__ cb, int cap) {
    constexpr int RUN = DEC_THREADS * PPT;
    if (mbits) {
        int pos = local_pos;
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            if ((mbits >> j) & 1u) {
                float4 P = __ldg(tab + ((code2[j >> 1] >> (16 * (j & 1))) & 0xFFFFu));
                s_out[pos] = s_x[col + j];
                s_out[RUN + pos] = yv;
                s_out[2 * RUN + pos] = P.x;
                s_out[3 * RUN + pos] = P.y;
                s_out[4 * RUN + pos] = P.z;
                pos++;
            }
        }
    }
    __syncthreads();
    const int n = min(total, max(cap - cta_base, 0));
#pragma unroll
    for (int pl = 0; pl < 5; pl++) {
        float* dst = cb + (size_t)pl * cap + cta_base;
        const float* src = s_out + pl * RUN;
        for (int i = threadIdx.x; i < n; i += DEC_THREADS) dst[i] = src[i];
    }
}

// -------------------------------------------------------------------------------------------------------------
// Fast path: vector loads, one CTA per run of 512*PPT pixels.  grid.x = B * ctas_per_crop.
// The compaction base of a CTA (masked pixels in the earlier runs of its crop) comes either from the other CTAs of a
// thread-block cluster through DSMEM (CLUSTER = true), or from re-counting the earlier part of the mask plane itself
// (CLUSTER = false: <= 7 extra 128-bit loads per thread, served by L2 because the owners of those runs stream the same
// lines; no cross-CTA dependency at all, so CTAs are scheduled and retire independently).
// Sign tests are one FSET (fp32) / HSET2 (bf16) per value producing an all-ones mask, and one LOP3 ors the plane's
// bit into codes kept two pixels per register: 2 issue slots per (pixel, plane) for fp32, 1 for bf16 -- the decode is
// as much issue-bound as HBM-bound (the first version spent ~7 slots per (pixel, plane)).
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gt0_mask_f32(uint32_t bits) {            // 0xFFFFFFFF if float(bits) > 0 (NaN, +-0 -> 0)
    uint32_t m;
    asm("set.gt.u32.f32 %0, %1, 0f00000000;" : "=r"(m) : "f"(__uint_as_float(bits)));
    return m;
}
__device__ __forceinline__ uint32_t gt0_mask_bf16x2(uint32_t two) {          // 0xFFFF per half that is > 0
    __nv_bfloat162 v;
    memcpy(&v, &two, 4);
    return __hgt2_mask(v, __nv_bfloat162(__float2bfloat16(0.f), __float2bfloat16(0.f)));
}
// ors bit `k` (a constant with the same 16-bit pattern in both halves) into the packed codes of the pixels of one
// 16-byte vector whose value is > 0.  code2[q] holds pixels 2q (low half) and 2q+1 (high half).
template <int DT> __device__ __forceinline__ void or_plane_bits(const uint4& v, uint32_t k, uint32_t* code2) {
    if (DT == ZP_DTYPE_F32) {
        code2[0] |= (gt0_mask_f32(v.x) & (k & 0xFFFFu)) | (gt0_mask_f32(v.y) & (k & 0xFFFF0000u));
        code2[1] |= (gt0_mask_f32(v.z) & (k & 0xFFFFu)) | (gt0_mask_f32(v.w) & (k & 0xFFFF0000u));
    } else {
        code2[0] |= gt0_mask_bf16x2(v.x) & k; code2[1] |= gt0_mask_bf16x2(v.y) & k;
        code2[2] |= gt0_mask_bf16x2(v.z) & k; code2[3] |= gt0_mask_bf16x2(v.w) & k;
    }
}

// FULL16: nb == 16 and the mask comes from the tensor (every shipped config at ignore_bit 0) -- compile-time plane count,
// no per-plane guards.
template <int DT, bool CLUSTER, bool FULL16>
__global__ void __launch_bounds__(DEC_THREADS, 2) zp_decode_cluster_kernel(DecodeArgs a, int ctas_per_crop) {
    constexpr int PPT = Px<DT>::N;
    constexpr int ESZ = Px<DT>::ESZ;
    const unsigned csize = (unsigned)ctas_per_crop;
    const int b = blockIdx.x / csize;
    const unsigned rank = blockIdx.x - b * csize;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.S, N = S * S;
    DEC_STAMP(0);

    __shared__ float s_x[1024], s_y[1024];     // remapped coordinates per column / row (S <= 1024 on this path)
    __shared__ int s_warp[DEC_WARPS];
    __shared__ int s_pre[DEC_WARPS];
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
    const char* crop = (const char*)a.logits + (size_t)b * a.sb * ESZ;
    const char* base = crop + ((size_t)row * a.sh + col) * ESZ;
    const size_t plane = (size_t)a.sc * ESZ;

    // ---- issue the mask load and the first half of the bit planes together (memory-level parallelism)
    uint32_t mbits = 0;
    uint4 v[8];
    const int nb = FULL16 ? 16 : a.nb;
    const bool own_mask = FULL16 || a.ext_mask == nullptr;
    int pre = 0;
    if (active) {
        uint4 mv = make_uint4(0, 0, 0, 0);
        if (own_mask) mv = zp_ldg_stream(base + (size_t)a.mask_ch * plane);
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (FULL16 || i < nb) v[i] = zp_ldg_stream(base + (size_t)(a.bit0_ch + i) * plane);
        if (own_mask) {
            mbits = positive_bits<DT>(mv);
        } else {
            const uint8_t* em = a.ext_mask + (size_t)b * N + p0;
#pragma unroll
            for (int j = 0; j < PPT; j++) mbits |= (uint32_t)(em[j] != 0) << j;
        }
    }
    if (!CLUSTER) {
        // masked pixels of runs 0..rank-1: thread t re-counts the pixels thread t of each earlier run owns (plain loads:
        // the lines stay in L1/L2 for the owners)
        for (unsigned r = 0; r < rank; r++) {
            const int q0 = ((int)r * DEC_THREADS + tid) * PPT;
            const int qrow = q0 / S, qcol = q0 - qrow * S;
            if (own_mask) {
                uint4 mv = __ldg((const uint4*)(crop + ((size_t)a.mask_ch * a.sc + (size_t)qrow * a.sh + qcol) * ESZ));
                pre += __popc(positive_bits<DT>(mv));
            } else {
                const uint8_t* em = a.ext_mask + (size_t)b * N + q0;
#pragma unroll
                for (int j = 0; j < PPT; j++) pre += em[j] != 0;
            }
        }
    }
    // ---- ranks of the masked pixels: thread -> warp -> CTA -> crop
    const int cnt = __popc(mbits);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (!CLUSTER) pre = __reduce_add_sync(0xffffffffu, pre);
    if (lane == 31) s_warp[warp] = incl;
    if (!CLUSTER && lane == 0) s_pre[warp] = pre;
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
        int pr = (!CLUSTER && lane < DEC_WARPS) ? s_pre[lane] : 0;
        pr = __reduce_add_sync(0xffffffffu, pr);
        if (lane == DEC_WARPS - 1) s_total = wi;
        if (lane == 0) s_pre[0] = pr;
    }
    __syncthreads();
    cg::cluster_group cluster = cg::this_cluster();
    if (CLUSTER && csize > 1) cluster.barrier_arrive();   // release s_total; the wait is after the code planes

    // ---- pack the code (two pixels per register, MSB-first at bit 15 - plane) while the exchange is in flight
    uint32_t code2[PPT / 2];
#pragma unroll
    for (int j = 0; j < PPT / 2; j++) code2[j] = 0;
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (FULL16 || i < nb) or_plane_bits<DT>(v[i], 0x80008000u >> i, code2);
        if (FULL16 || nb > 8) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (FULL16 || 8 + i < nb) v[i] = zp_ldg_stream(base + (size_t)(a.bit0_ch + 8 + i) * plane);
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (FULL16 || 8 + i < nb) or_plane_bits<DT>(v[i], 0x00800080u >> i, code2);
        }
        const int sh = 16 - nb;                                 // planes were placed as if nb == 16
#pragma unroll
        for (int j = 0; j < PPT / 2; j++) code2[j] = (code2[j] >> sh) & (0xFFFFu >> sh) * 0x00010001u;
        if (a.codes) {
            uint16_t* cp = a.codes + (size_t)b * N + p0;
            if (PPT == 4) *(uint2*)cp = make_uint2(code2[0], code2[1]);
            else *(uint4*)cp = make_uint4(code2[0], code2[1], code2[2 % (PPT / 2)], code2[3 % (PPT / 2)]);
        }
    }

    DEC_STAMP(1);
    // ---- crop-wide exclusive prefix of the CTA totals
    int cta_base = 0;
    if (CLUSTER) {
        if (csize > 1) {
            cluster.barrier_wait();
            for (unsigned r = 0; r < rank; r++) cta_base += *cluster.map_shared_rank(&s_total, r);
            cluster.barrier_arrive();                         // peers may retire once everybody has read
        }
    } else {
        cta_base = s_pre[0];
    }
    if (rank == csize - 1 && tid == 0) a.counts[b] = cta_base + s_total;

    // ---- gather the 3D points and emit the run in row-major order (staged through shared memory)
    {
        extern __shared__ __align__(16) float s_out[];
        const int obj = zp_obj_slot(a.obj_ids, a.obj_default, b);
        emit_staged<PPT>(s_out, mbits, code2, s_warp[warp] + incl - cnt, a.tables[obj], s_x, col, s_y[row], s_total, cta_base,
                         a.corr + (size_t)b * 5 * a.cap, a.cap);
    }
    if (CLUSTER && csize > 1) cluster.barrier_wait();
    DEC_STAMP(3);
}

// -------------------------------------------------------------------------------------------------------------
// Default path: the fused kernel above, but a CTA walks `rpc` CONSECUTIVE runs of its crop and issues the mask + first
// eight plane loads of run i+1 before it gathers / stages / writes run i.  Per-CTA %globaltimer stamps of the single-run
// kernel (tools/dbg_decode_ctas.py, 64 crops) showed why it stops at 39 % of HBM peak: all 296 resident CTAs start
// together, so they all read (8 us, HBM saturated), then all emit (3.6 us, HBM idle), then the second wave does the same:
// the phases are in lock-step across the whole chip.  With several runs per CTA the grid fits in one wave (64 crops:
// 256 CTAs of 2 runs) and every emit phase is covered by the next run's loads; the base of a later run is the running
// total, so only the first run of a CTA re-counts earlier mask pixels.
// -------------------------------------------------------------------------------------------------------------
template <int DT, bool FULL16>
__global__ void __launch_bounds__(DEC_THREADS, 2) zp_decode_stream_kernel(DecodeArgs a, int ctas_per_crop, int rpc, int runs_per_crop) {
    constexpr int PPT = Px<DT>::N;
    constexpr int ESZ = Px<DT>::ESZ;
    const int b = blockIdx.x / ctas_per_crop;
    const int rank0 = (blockIdx.x - b * ctas_per_crop) * rpc;            // first run of this CTA
    const int n_runs = min(rpc, runs_per_crop - rank0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.S, N = S * S;
    DEC_STAMP(0);
    __shared__ float s_x[1024], s_y[1024];
    __shared__ int s_warp[2][DEC_WARPS];
    __shared__ int s_pre[DEC_WARPS];
    __shared__ int s_total[2];
    extern __shared__ __align__(16) float s_out[];
    const char* crop = (const char*)a.logits + (size_t)b * a.sb * ESZ;
    const size_t plane = (size_t)a.sc * ESZ;
    const int nb = FULL16 ? 16 : a.nb;
    const bool own_mask = FULL16 || a.ext_mask == nullptr;
    const int obj = zp_obj_slot(a.obj_ids, a.obj_default, b);
    const float4* tab = a.tables[obj];
    float* cb = a.corr + (size_t)b * 5 * a.cap;

    uint4 mv = make_uint4(0, 0, 0, 0), v[8];
    auto issue_first = [&](int rank) {                       // mask + planes 0..7 of run `rank`
        const int p0 = (rank * DEC_THREADS + tid) * PPT;
        if (p0 < N) {
            const int row = p0 / S, col = p0 - row * S;
            const char* base = crop + ((size_t)row * a.sh + col) * ESZ;
            if (own_mask) mv = zp_ldg_stream(base + (size_t)a.mask_ch * plane);
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (FULL16 || i < nb) v[i] = zp_ldg_stream(base + (size_t)(a.bit0_ch + i) * plane);
        }
    };
    issue_first(rank0);
    // masked pixels of the runs before this CTA's first one (plain loads: L2 hits, the owners stream the same lines)
    int pre = 0;
    for (int r = 0; r < rank0; r++) {
        const int q0 = (r * DEC_THREADS + tid) * PPT;
        const int qrow = q0 / S, qcol = q0 - qrow * S;
        if (own_mask) {
            uint4 m = __ldg((const uint4*)(crop + ((size_t)a.mask_ch * a.sc + (size_t)qrow * a.sh + qcol) * ESZ));
            pre += __popc(positive_bits<DT>(m));
        } else {
            const uint8_t* em = a.ext_mask + (size_t)b * N + q0;
#pragma unroll
            for (int j = 0; j < PPT; j++) pre += em[j] != 0;
        }
    }
    {   // per-crop coordinate LUTs (after the loads are in flight)
        const double* bb = a.bbox + 4 * (size_t)b;
        double x0 = bb[0], y0 = bb[1], w = bb[2], h = bb[3];
        for (int i = tid; i < S; i += DEC_THREADS) {
            s_x[i] = remap_coord(w, x0, S, i);
            s_y[i] = remap_coord(h, y0, S, i);
        }
    }
    pre = __reduce_add_sync(0xffffffffu, pre);
    if (lane == 0) s_pre[warp] = pre;
    int cta_base = -1;                                       // known after the first barrier

    for (int it = 0; it < n_runs; it++) {
        const int rank = rank0 + it;
        const int p0 = (rank * DEC_THREADS + tid) * PPT;
        const bool active = p0 < N;
        const int row = active ? p0 / S : 0, col = active ? p0 - row * S : 0;
        const char* base = crop + ((size_t)row * a.sh + col) * ESZ;
        uint32_t mbits = 0;
        if (active) {
            if (own_mask) mbits = positive_bits<DT>(mv);
            else {
                const uint8_t* em = a.ext_mask + (size_t)b * N + p0;
#pragma unroll
                for (int j = 0; j < PPT; j++) mbits |= (uint32_t)(em[j] != 0) << j;
            }
        }
        const int cnt = __popc(mbits);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[it & 1][warp] = incl;
        // ---- pack the code while the other warps arrive
        uint32_t code2[PPT / 2];
#pragma unroll
        for (int j = 0; j < PPT / 2; j++) code2[j] = 0;
        if (active) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (FULL16 || i < nb) or_plane_bits<DT>(v[i], 0x80008000u >> i, code2);
            if (FULL16 || nb > 8) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (FULL16 || 8 + i < nb) v[i] = zp_ldg_stream(base + (size_t)(a.bit0_ch + 8 + i) * plane);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (FULL16 || 8 + i < nb) or_plane_bits<DT>(v[i], 0x00800080u >> i, code2);
            }
            const int sh = 16 - nb;
#pragma unroll
            for (int j = 0; j < PPT / 2; j++) code2[j] = (code2[j] >> sh) & (0xFFFFu >> sh) * 0x00010001u;
            if (a.codes) {
                uint16_t* cp = a.codes + (size_t)b * N + p0;
                if (PPT == 4) *(uint2*)cp = make_uint2(code2[0], code2[1]);
                else *(uint4*)cp = make_uint4(code2[0], code2[1], code2[2 % (PPT / 2)], code2[3 % (PPT / 2)]);
            }
        }
        if (it + 1 < n_runs) issue_first(rank + 1);          // next run's loads fly during this run's emit
        __syncthreads();                                      // warp totals (and, first time, s_pre and the LUTs) visible
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < DEC_WARPS; w++) {
            const int t = s_warp[it & 1][w];
            wbase += w < warp ? t : 0;
            total += t;
        }
        if (cta_base < 0) {
            cta_base = 0;
#pragma unroll
            for (int w = 0; w < DEC_WARPS; w++) cta_base += s_pre[w];
        }
        if (it == 0) DEC_STAMP(1);
        emit_staged<PPT>(s_out, mbits, code2, wbase + incl - cnt, tab, s_x, col, s_y[row], total, cta_base, cb, a.cap);
        cta_base += total;
        __syncthreads();                                      // s_out is reused by the next run
    }
    if (rank0 + n_runs == runs_per_crop && tid == 0) a.counts[b] = cta_base;
    DEC_STAMP(3);
}

template <int DT>
static int launch_stream(zp_ctx* ctx, const DecodeArgs& a, int runs, cudaStream_t st) {
    const int smem = 5 * DEC_THREADS * Px<DT>::N * (int)sizeof(float);
    static bool attr_set_dev[ZP_MAX_DEVICES] = {};          // cudaFuncSetAttribute is per device
    bool& attr_set = attr_set_dev[ctx->device % ZP_MAX_DEVICES];
    if (!attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_stream_kernel<DT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_stream_kernel<DT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    // runs per CTA: the smallest count that brings the grid into one wave of 2 CTAs per SM, but at most 4 (measured at
    // 1024 crops: 1 run 71 %, 2 runs 75.5 %, 4 runs 77.8 %, 8 runs 68 % of HBM peak -- long CTAs lose to the tail)
    const int slots = 2 * ctx->sm_count;
    int rpc = ctx->decode_rpc > 0 ? ctx->decode_rpc : 1;
    if (ctx->decode_rpc <= 0)
        while (rpc < runs && rpc < 4 && 2 * (rpc + 1) <= runs && (long long)a.B * ((runs + rpc - 1) / rpc) > slots) rpc++;   // >= 2 CTAs per crop
    if (rpc > runs) rpc = runs;
    const int cpc = (runs + rpc - 1) / rpc;
    const bool full16 = a.nb == 16 && a.ext_mask == nullptr;
    ZP_TIME_BEGIN(ctx, st);
    if (full16) zp_decode_stream_kernel<DT, true><<<(unsigned)(a.B * cpc), DEC_THREADS, smem, st>>>(a, cpc, rpc, runs);
    else zp_decode_stream_kernel<DT, false><<<(unsigned)(a.B * cpc), DEC_THREADS, smem, st>>>(a, cpc, rpc, runs);
    ZP_CHECK_LAUNCH(ctx, "zp_decode_stream_kernel");
    return 0;
}

// -------------------------------------------------------------------------------------------------------------
// Two-kernel path (default).  The fused kernels above and below are bound by the LIFETIME of a CTA, not by bytes: load
// -> block scan -> load -> table gather -> scattered stores are five dependent memory round trips during most of which
// the CTA has no plane loads in flight (bf16 input, half the bytes, takes the same time as fp32).  Splitting the work
// removes every dependency from the part that moves 97 % of the bytes:
//   zp_decode_planes_kernel  pure stream: a thread issues the 128-bit loads of ALL planes of its PPT pixels at once,
//                            packs the codes (FSET/HSET2 + LOP3) and writes 2 B/pixel of codes + 1 bit/pixel of mask
//                            (warp ballots).  No shared memory, no barrier, one memory round trip per thread.
//   zp_decode_emit_kernel    per run of 512*PPT pixels: ranks from the mask bits (the base of a run = popcount of the
//                            earlier runs' ballot words, <= 2 KB), table gather, emit.  Touches 2.1 + 15 B/pixel.
// -------------------------------------------------------------------------------------------------------------
constexpr int PL_THREADS = 128;

template <int DT>
__global__ void __launch_bounds__(PL_THREADS, 5) zp_decode_planes_kernel(DecodeArgs a, uint16_t* __restrict__ codes,
                                                                          uint32_t* __restrict__ maskw, int segs_per_crop) {
    constexpr int PPT = Px<DT>::N;
    constexpr int ESZ = Px<DT>::ESZ;
    const int S = a.S, N = S * S, nb = a.nb;
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * (PL_THREADS / 32) + (threadIdx.x >> 5);        // 32*PPT consecutive pixels of one crop
    const int b = seg / segs_per_crop, sg = seg - b * segs_per_crop;
    if (b >= a.B) return;
    const int p0 = (sg * 32 + lane) * PPT;
    const bool active = p0 < N;
    uint32_t mbits = 0;
    if (active) {
        const int row = p0 / S, col = p0 - row * S;
        const char* base = (const char*)a.logits + ((size_t)b * a.sb + (size_t)row * a.sh + col) * ESZ;
        const size_t plane = (size_t)a.sc * ESZ;
        uint4 mv = make_uint4(0, 0, 0, 0), v[16];
        if (a.ext_mask == nullptr) mv = zp_ldg_stream(base + (size_t)a.mask_ch * plane);
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i < nb) v[i] = zp_ldg_stream(base + (size_t)(a.bit0_ch + i) * plane);
        if (a.ext_mask == nullptr) {
            mbits = positive_bits<DT>(mv);
        } else {
            const uint8_t* em = a.ext_mask + (size_t)b * N + p0;
#pragma unroll
            for (int j = 0; j < PPT; j++) mbits |= (uint32_t)(em[j] != 0) << j;
        }
        uint32_t code2[PPT / 2];
#pragma unroll
        for (int j = 0; j < PPT / 2; j++) code2[j] = 0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i < nb) or_plane_bits<DT>(v[i], 0x80008000u >> i, code2);
        const int sh = 16 - nb;                                     // planes were placed as if nb == 16
#pragma unroll
        for (int j = 0; j < PPT / 2; j++) code2[j] = (code2[j] >> sh) & (0xFFFFu >> sh) * 0x00010001u;
        uint16_t* cp = codes + (size_t)b * N + p0;
        if (PPT == 4) *(uint2*)cp = make_uint2(code2[0], code2[1]);
        else *(uint4*)cp = make_uint4(code2[0], code2[1], code2[2 % (PPT / 2)], code2[3 % (PPT / 2)]);
    }
    // mask bits of the segment: word j, bit i = pixel PPT*i + j masked
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < PPT; j++) {
        uint32_t bal = __ballot_sync(0xffffffffu, (mbits >> j) & 1u);
        if (lane == j) mine = bal;
    }
    if (lane < PPT) maskw[((size_t)b * segs_per_crop + sg) * PPT + lane] = mine;
}

template <int PPT>
__global__ void __launch_bounds__(DEC_THREADS, 3) zp_decode_emit_kernel(DecodeArgs a, const uint16_t* __restrict__ codes,
                                                                        const uint32_t* __restrict__ maskw,
                                                                        int segs_per_crop, int runs_per_crop) {
    const int b = blockIdx.x / runs_per_crop, run = blockIdx.x - b * runs_per_crop;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.S, N = S * S;
    __shared__ float s_x[1024], s_y[1024];
    __shared__ int s_warp[DEC_WARPS];
    __shared__ int s_pre[DEC_WARPS];
    __shared__ int s_total;
    {
        const double* bb = a.bbox + 4 * (size_t)b;
        double x0 = bb[0], y0 = bb[1], w = bb[2], h = bb[3];
        for (int i = tid; i < S; i += DEC_THREADS) {
            s_x[i] = remap_coord(w, x0, S, i);
            s_y[i] = remap_coord(h, y0, S, i);
        }
    }
    const uint32_t* mw = maskw + (size_t)b * segs_per_crop * PPT;
    const int sg = run * DEC_WARPS + warp;                           // this warp's segment
    const int p0 = (sg * 32 + lane) * PPT;
    const bool active = p0 < N;
    uint32_t mbits = 0;
    uint32_t c2[PPT / 2];
#pragma unroll
    for (int j = 0; j < PPT / 2; j++) c2[j] = 0;
    if (sg < segs_per_crop) {
#pragma unroll
        for (int j = 0; j < PPT; j++) mbits |= ((__ldg(mw + (size_t)sg * PPT + j) >> lane) & 1u) << j;
    }
    if (active && mbits) {
        const uint16_t* cp = codes + (size_t)b * N + p0;
        if (PPT == 4) { uint2 t = __ldg((const uint2*)cp); c2[0] = t.x; c2[1] = t.y; }
        else { uint4 t = __ldg((const uint4*)cp); c2[0] = t.x; c2[1] = t.y; c2[2 % (PPT / 2)] = t.z; c2[3 % (PPT / 2)] = t.w; }
    }
    // masked pixels of the earlier runs: popcount of their ballot words
    int pre = 0;
    for (int i = tid; i < run * DEC_WARPS * PPT; i += DEC_THREADS) pre += __popc(__ldg(mw + i));
    const int cnt = __popc(mbits);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    pre = __reduce_add_sync(0xffffffffu, pre);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_pre[warp] = pre;
    __syncthreads();
    if (warp == 0) {
        int w = lane < DEC_WARPS ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < DEC_WARPS; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < DEC_WARPS) s_warp[lane] = wi - w;
        int pr = lane < DEC_WARPS ? s_pre[lane] : 0;
        pr = __reduce_add_sync(0xffffffffu, pr);
        if (lane == DEC_WARPS - 1) s_total = wi;
        if (lane == 0) s_pre[0] = pr;
    }
    __syncthreads();
    const int cta_base = s_pre[0];
    if (run == runs_per_crop - 1 && tid == 0) a.counts[b] = cta_base + s_total;
    {
        extern __shared__ __align__(16) float s_out[];
        const int obj = zp_obj_slot(a.obj_ids, a.obj_default, b);
        const int row = active ? p0 / S : 0, col = active ? p0 - row * S : 0;
        emit_staged<PPT>(s_out, mbits, c2, s_warp[warp] + incl - cnt, a.tables[obj], s_x, col, s_y[row], s_total, cta_base,
                         a.corr + (size_t)b * 5 * a.cap, a.cap);
    }
}

static int dws_reserve(zp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->dws_bytes) return 0;
    ZP_CUDA(ctx, cudaDeviceSynchronize());
    zp_drop_graphs(ctx);
    if (ctx->dws) cudaFree(ctx->dws);
    ctx->dws = nullptr; ctx->dws_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    ZP_CUDA(ctx, cudaMalloc(&ctx->dws, want));
    ctx->dws_bytes = want;
    return 0;
}

template <int DT>
static int launch_split(zp_ctx* ctx, const DecodeArgs& a, cudaStream_t st) {
    constexpr int PPT = Px<DT>::N;
    const int N = a.S * a.S;
    const int segs = (N + 32 * PPT - 1) / (32 * PPT);
    const int runs = (segs + DEC_WARPS - 1) / DEC_WARPS;
    const size_t code_bytes = a.codes ? 0 : (((size_t)a.B * N * sizeof(uint16_t) + 255) & ~(size_t)255);
    const size_t mask_bytes = (size_t)a.B * segs * PPT * sizeof(uint32_t);
    if (dws_reserve(ctx, code_bytes + mask_bytes)) return -2;
    uint16_t* codes = a.codes ? a.codes : (uint16_t*)ctx->dws;
    uint32_t* maskw = (uint32_t*)((char*)ctx->dws + code_bytes);
    const long long total_segs = (long long)a.B * segs;
    const unsigned grid1 = (unsigned)((total_segs + PL_THREADS / 32 - 1) / (PL_THREADS / 32));
    zp_decode_planes_kernel<DT><<<grid1, PL_THREADS, 0, st>>>(a, codes, maskw, segs);
    ZP_CHECK_LAUNCH(ctx, "zp_decode_planes_kernel");
    const int smem = 5 * DEC_THREADS * PPT * (int)sizeof(float);
    static bool attr_set_dev[ZP_MAX_DEVICES] = {};          // cudaFuncSetAttribute is per device
    bool& attr_set = attr_set_dev[ctx->device % ZP_MAX_DEVICES];
    if (!attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_emit_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    zp_decode_emit_kernel<PPT><<<(unsigned)(a.B * runs), DEC_THREADS, smem, st>>>(a, codes, maskw, segs, runs);
    ZP_CHECK_LAUNCH(ctx, "zp_decode_emit_kernel");
    return 0;
}

// Second half of the two-kernel path on its own, for producers that already hold 2 B/pixel codes + mask ballot words
// in the fp32 flavour's layout (segments of 128 pixels, word j bit i = pixel 4i + j): the fused network head
// (zp_head.cu) writes exactly that from its tensor-core epilogue.
int zp_launch_emit_codes(zp_ctx* ctx, int B, int S, const double* bbox, const int32_t* obj_ids, int obj_default,
                         const uint16_t* codes, const uint32_t* maskw, float* corr, int cap, int32_t* counts, cudaStream_t st) {
    constexpr int PPT = 4;
    DecodeArgs a{};
    a.B = B; a.S = S; a.bbox = bbox; a.obj_ids = obj_ids; a.obj_default = obj_default;
    a.tables = (const float4* const*)ctx->d_table_ptrs;
    a.corr = corr; a.cap = cap; a.counts = counts;
    const int N = S * S;
    const int segs = (N + 32 * PPT - 1) / (32 * PPT);
    const int runs = (segs + DEC_WARPS - 1) / DEC_WARPS;
    if (S % PPT != 0 || S > 1024 || runs > 64) ZP_FAIL(ctx, -1, "emit from codes: crop size %d not supported", S);
    const int smem = 5 * DEC_THREADS * PPT * (int)sizeof(float);
    static bool attr_set_dev[ZP_MAX_DEVICES] = {};          // cudaFuncSetAttribute is per device
    bool& attr_set = attr_set_dev[ctx->device % ZP_MAX_DEVICES];
    if (!attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_emit_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    ZP_TIME_BEGIN(ctx, st);
    zp_decode_emit_kernel<PPT><<<(unsigned)(B * runs), DEC_THREADS, smem, st>>>(a, codes, maskw, segs, runs);
    ZP_CHECK_LAUNCH(ctx, "zp_decode_emit_kernel");
    return 0;
}

// -------------------------------------------------------------------------------------------------------------
// CE heads (ablation configs, class_base = divided_num_each_interation > 2 or CE loss; common_ops.py:21-30): the code
// logits are n_digits groups of `base` consecutive channels; the digit is the FIRST maximum of the float32 softmax of its
// group (np.argmax of torch.softmax), the class id the base-`base` number with digit 0 most significant
// (class_id_encoder_decoder.py:17-28).  One thread per pixel, one 128-pixel segment per CTA; writes the same 2 B/pixel
// codes + mask ballot words as the binary plane kernel, so zp_decode_emit_kernel finishes the job.  Not a hot path.
// -------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(128) zp_decode_ce_kernel(DecodeArgs a, int base, int n_digits, uint16_t* __restrict__ codes,
                                                            uint32_t* __restrict__ maskw, int segs_per_crop) {
    const int S = a.S, N = S * S;
    const int b = blockIdx.x / segs_per_crop, sg = blockIdx.x - b * segs_per_crop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = sg * 128 + threadIdx.x;
    bool masked = false;
    if (p < N) {
        const int row = p / S, col = p - row * S;
        auto at = [&](int ch) -> float {
            const size_t off = (size_t)b * a.sb + (size_t)ch * a.sc + (size_t)row * a.sh + (size_t)col * a.sw;
            if (DT == ZP_DTYPE_F32) return ((const float*)a.logits)[off];
            return __bfloat162float(((const __nv_bfloat16*)a.logits)[off]);
        };
        masked = a.ext_mask ? a.ext_mask[(size_t)b * N + p] != 0 : at(a.mask_ch) > 0.0f;
        uint32_t id = 0;
        for (int d = 0; d < n_digits; d++) {
            const int c0 = a.bit0_ch + d * base;
            float m = at(c0);
            for (int k = 1; k < base; k++) m = fmaxf(m, at(c0 + k));
            float sum = 0.f;
            for (int k = 0; k < base; k++) sum = __fadd_rn(sum, expf(__fsub_rn(at(c0 + k), m)));
            int best = 0;
            float pb = -1.f;
            for (int k = 0; k < base; k++) {
                const float pk = __fdiv_rn(expf(__fsub_rn(at(c0 + k), m)), sum);
                if (pk > pb) { pb = pk; best = k; }          // strictly greater: first maximum, as numpy.argmax
            }
            id = id * (uint32_t)base + (uint32_t)best;
        }
        codes[(size_t)b * N + p] = (uint16_t)id;
    }
    // mask ballot: pixel 4i + j of the segment -> bit i of word j (byte `warp` of each word belongs to this warp)
    const uint32_t bal = __ballot_sync(0xffffffffu, masked);
    if (lane < 4) {
        uint32_t byte = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) byte |= ((bal >> (4 * k + lane)) & 1u) << k;
        ((uint8_t*)maskw)[((size_t)(b * segs_per_crop + sg) * 4 + lane) * 4 + warp] = (uint8_t)byte;
    }
}

int zp_launch_decode_ce(zp_ctx* ctx, const void* logits, int dtype, int B, int S, const int64_t strides[4], int mask_ch,
                        int digit0_ch, int base, int n_digits, const uint8_t* ext_mask, const double* bbox,
                        const int32_t* obj_ids, int obj_default, uint16_t* codes, float* corr, int cap, int32_t* counts,
                        cudaStream_t st) {
    DecodeArgs a{};
    a.logits = logits; a.sb = strides[0]; a.sc = strides[1]; a.sh = strides[2]; a.sw = strides[3];
    a.B = B; a.S = S; a.mask_ch = mask_ch; a.bit0_ch = digit0_ch; a.ext_mask = ext_mask;
    const int N = S * S;
    const int segs = (N + 127) / 128;
    const size_t code_bytes = codes ? 0 : (((size_t)B * N * sizeof(uint16_t) + 255) & ~(size_t)255);
    const size_t mask_bytes = (size_t)B * segs * 4 * sizeof(uint32_t);
    if (dws_reserve(ctx, code_bytes + mask_bytes)) return -2;
    uint16_t* d_codes = codes ? codes : (uint16_t*)ctx->dws;
    uint32_t* maskw = (uint32_t*)((char*)ctx->dws + code_bytes);
    if (dtype == ZP_DTYPE_F32) zp_decode_ce_kernel<ZP_DTYPE_F32><<<(unsigned)(B * segs), 128, 0, st>>>(a, base, n_digits, d_codes, maskw, segs);
    else zp_decode_ce_kernel<ZP_DTYPE_BF16><<<(unsigned)(B * segs), 128, 0, st>>>(a, base, n_digits, d_codes, maskw, segs);
    ZP_CHECK_LAUNCH(ctx, "zp_decode_ce_kernel");
    return zp_launch_emit_codes(ctx, B, S, bbox, obj_ids, obj_default, d_codes, maskw, corr, cap, counts, st);
}

// -------------------------------------------------------------------------------------------------------------
// Streaming path (contiguous planes): a producer warp pulls the logits of a crop part through a shared-memory ring of
// 8 KB plane segments with 1-D TMA bulk copies (cp.async.bulk + full/empty mbarriers, SASS UBLKCP), so the memory
// system always has up to 12 x 8 KB per CTA in flight while the four consumer warps pack codes, rank, gather and
// store -- the register-staged kernel above has no loads in flight during its scan / gather / store phases and tops
// out at ~36 % (64 crops) .. 56 % (1024 crops) of HBM peak.  (Copies are 8 KB because the TMA unit spends a fixed
// ~250 cycles per bulk copy: a first version with 2 KB copies reached only 2.4 TB/s.)
// Work: a crop is split into `parts` contiguous pixel ranges, one CTA each, launched as a cluster of `parts` CTAs; a
// CTA walks its range in blocks of 8 KB / sizeof(logit) pixels, plane by plane, keeping the partial codes of its
// pixels in registers, with a running output offset.  The only cross-CTA dependency -- the number of masked pixels in
// the earlier parts -- is resolved once, up front, by counting the own mask range (plain vector loads) and exchanging
// the totals through DSMEM while the ring fills.
// -------------------------------------------------------------------------------------------------------------
constexpr int TMA_CONSUMERS = 128;
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;  // + one producer warp
constexpr int TMA_SLOT_BYTES = 8192;
constexpr int TMA_SLOTS = 12;
constexpr int TMA_MAX_PARTS = 8;

__device__ __forceinline__ uint32_t dsm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dec_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dsm_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dec_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsm_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dec_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dsm_u32(bar)) : "memory");
}
__device__ __forceinline__ void dec_mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "DEC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DEC_DONE;\n\t"
        "bra DEC_WAIT;\n\t"
        "DEC_DONE:\n\t}" ::"r"(dsm_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void dec_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dsm_u32(dst)), "l"(src), "r"(bytes), "r"(dsm_u32(bar)) : "memory");
}

// 4 consecutive pixels (index q4 = pixel / 4 within the slot) -> 4 sign bits
template <int DT> __device__ __forceinline__ uint32_t slot_bits4(const unsigned char* slot, int q4) {
    if (DT == ZP_DTYPE_F32) {
        uint4 v = *(const uint4*)(slot + (size_t)q4 * 16);
        return (uint32_t)pos_f32(v.x) | ((uint32_t)pos_f32(v.y) << 1) | ((uint32_t)pos_f32(v.z) << 2) |
               ((uint32_t)pos_f32(v.w) << 3);
    } else {
        uint2 v = *(const uint2*)(slot + (size_t)q4 * 8);
        return (uint32_t)pos_bf16(v.x) | ((uint32_t)pos_bf16(v.x >> 16) << 1) | ((uint32_t)pos_bf16(v.y) << 2) |
               ((uint32_t)pos_bf16(v.y >> 16) << 3);
    }
}

template <int DT>
__global__ void __launch_bounds__(TMA_THREADS, 2) zp_decode_tma_kernel(DecodeArgs a, int parts, int part_px) {
    constexpr int ESZ = Px<DT>::ESZ;
    constexpr int BLK = TMA_SLOT_BYTES / ESZ;                         // pixels per block (= per ring slot)
    constexpr int G = BLK / (TMA_CONSUMERS * 4);                      // 4-pixel groups per consumer thread and block
    extern __shared__ __align__(128) unsigned char s_ring[];          // [TMA_SLOTS][TMA_SLOT_BYTES]
    __shared__ __align__(8) uint64_t s_full[TMA_SLOTS], s_empty[TMA_SLOTS];
    __shared__ unsigned long long s_warp[2][TMA_CONSUMERS / 32];
    __shared__ int s_total;                                           // masked pixels of this part (read by peers)
    __shared__ int s_red[TMA_THREADS / 32];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = parts > 1 ? (int)cluster.block_rank() : 0;
    const int b = blockIdx.x / parts;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.S, N = S * S, nb = a.nb;
    const bool own_mask = a.ext_mask == nullptr;
    const int n_planes = nb + (own_mask ? 1 : 0);                     // planes per block; the mask plane (if any) first
    const int px0 = min(N, rank * part_px), px1 = min(N, px0 + part_px);
    const int n_blocks = (px1 - px0 + BLK - 1) / BLK;
    const int n_copies = n_blocks * n_planes;
    const char* crop = (const char*)a.logits + (size_t)b * a.sb * ESZ;
    const bool producer = warp == TMA_CONSUMERS / 32;

    auto issue = [&](int q) {                                         // copy q = (block, plane) into slot q % TMA_SLOTS
        const int blk = q / n_planes, pl = q - blk * n_planes, slot = q % TMA_SLOTS;
        const int p = px0 + blk * BLK;
        const uint32_t bytes = (uint32_t)(min(BLK, px1 - p) * ESZ);
        const int ch = own_mask ? (pl == 0 ? a.mask_ch : a.bit0_ch + pl - 1) : a.bit0_ch + pl;
        dec_mbar_expect_tx(&s_full[slot], bytes);
        dec_tma_load_1d(s_ring + (size_t)slot * TMA_SLOT_BYTES, crop + ((size_t)ch * a.sc + p) * ESZ, bytes, &s_full[slot]);
    };
    if (tid == TMA_CONSUMERS) {
        for (int s = 0; s < TMA_SLOTS; s++) { dec_mbar_init(&s_full[s], 1); dec_mbar_init(&s_empty[s], TMA_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int q = 0; q < TMA_SLOTS && q < n_copies; q++) issue(q);
    }
    // ---- masked pixels in the earlier parts of this crop (only when the crop is split)
    int running = 0;
    if (parts > 1) {
        int cnt = 0;
        if (!producer)
            for (int p = px0 + tid * 4; p < px1; p += TMA_CONSUMERS * 4) {
                if (own_mask) {
                    const char* src = crop + ((size_t)a.mask_ch * a.sc + p) * ESZ;
                    if (DT == ZP_DTYPE_F32) {
                        uint4 v = zp_ldg_stream(src);
                        cnt += (int)pos_f32(v.x) + (int)pos_f32(v.y) + (int)pos_f32(v.z) + (int)pos_f32(v.w);
                    } else {
                        uint2 v = *(const uint2*)src;
                        cnt += (int)pos_bf16(v.x) + (int)pos_bf16(v.x >> 16) + (int)pos_bf16(v.y) + (int)pos_bf16(v.y >> 16);
                    }
                } else {
                    uchar4 m = *(const uchar4*)(a.ext_mask + (size_t)b * N + p);
                    cnt += (m.x != 0) + (m.y != 0) + (m.z != 0) + (m.w != 0);
                }
            }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) s_red[warp] = cnt;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < TMA_THREADS / 32; w++) t += s_red[w]; s_total = t; }
        cluster.sync();
        for (int r = 0; r < rank; r++) running += *cluster.map_shared_rank(&s_total, r);
        cluster.sync();                                               // peers may exit once everybody has read
    } else {
        __syncthreads();                                              // mbarrier init visible to the waiting threads
    }
    if (producer) {
        if (lane == 0)
            for (int q = TMA_SLOTS; q < n_copies; q++) {
                const int slot = q % TMA_SLOTS;
                dec_mbar_wait(&s_empty[slot], (uint32_t)(q / TMA_SLOTS - 1) & 1u);
                issue(q);
            }
        return;
    }
    // ---- consumers
    const double* bb = a.bbox + 4 * (size_t)b;
    const double x0 = bb[0], y0 = bb[1], rx = bb[2] / (double)S, ry = bb[3] / (double)S;
    const int obj = zp_obj_slot(a.obj_ids, a.obj_default, b);
    const float4* tab = a.tables[obj];
    float* cb = a.corr + (size_t)b * 5 * a.cap;
    const size_t cap = (size_t)a.cap;
    int q = 0;
    for (int blk = 0; blk < n_blocks; blk++) {
        const int pb = px0 + blk * BLK;                               // first pixel of the block
        uint32_t mbits[G], code[G][4];
#pragma unroll
        for (int g = 0; g < G; g++) { mbits[g] = 0; code[g][0] = code[g][1] = code[g][2] = code[g][3] = 0; }
        if (!own_mask) {
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int p = pb + (g * TMA_CONSUMERS + tid) * 4;
                if (p < px1) {
                    uchar4 m = *(const uchar4*)(a.ext_mask + (size_t)b * N + p);
                    mbits[g] = (uint32_t)(m.x != 0) | ((uint32_t)(m.y != 0) << 1) | ((uint32_t)(m.z != 0) << 2) | ((uint32_t)(m.w != 0) << 3);
                }
            }
        }
        for (int pl = 0; pl < n_planes; pl++, q++) {
            const int slot = q % TMA_SLOTS;
            dec_mbar_wait(&s_full[slot], (uint32_t)(q / TMA_SLOTS) & 1u);
            const unsigned char* sl = s_ring + (size_t)slot * TMA_SLOT_BYTES;
            uint32_t pbits[G];
#pragma unroll
            for (int g = 0; g < G; g++) pbits[g] = slot_bits4<DT>(sl, g * TMA_CONSUMERS + tid);
            __syncwarp();
            if (lane == 0) dec_mbar_arrive(&s_empty[slot]);           // this warp is done with the slot
            if (own_mask && pl == 0) {
#pragma unroll
                for (int g = 0; g < G; g++) mbits[g] = (pb + (g * TMA_CONSUMERS + tid) * 4 < px1) ? pbits[g] : 0u;
            } else {
                const int sh = nb - 1 - (own_mask ? pl - 1 : pl);
#pragma unroll
                for (int g = 0; g < G; g++) {
                    code[g][0] |= (pbits[g] & 1u) << sh; code[g][1] |= ((pbits[g] >> 1) & 1u) << sh;
                    code[g][2] |= ((pbits[g] >> 2) & 1u) << sh; code[g][3] |= ((pbits[g] >> 3) & 1u) << sh;
                }
            }
        }
        if (a.codes) {
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int p = pb + (g * TMA_CONSUMERS + tid) * 4;
                if (p < px1) *(uint2*)(a.codes + (size_t)b * N + p) = make_uint2(code[g][0] | (code[g][1] << 16), code[g][2] | (code[g][3] << 16));
            }
        }
        // ---- ranks: the G per-group counts (<= 4 per thread, <= 128 per warp) ride in the bytes of one 64-bit word
        unsigned long long pk = 0;
#pragma unroll
        for (int g = 0; g < G; g++) pk |= (unsigned long long)__popc(mbits[g]) << (8 * g);
        unsigned long long incl = pk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[blk & 1][warp] = incl;
        asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
        unsigned long long wv[TMA_CONSUMERS / 32];
#pragma unroll
        for (int w = 0; w < TMA_CONSUMERS / 32; w++) wv[w] = s_warp[blk & 1][w];
        int gbase = running;
#pragma unroll
        for (int g = 0; g < G; g++) {
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < TMA_CONSUMERS / 32; w++) {
                const int v = (int)((wv[w] >> (8 * g)) & 0xffu);
                before += w < warp ? v : 0;
                total += v;
            }
            const int cnt = (int)((pk >> (8 * g)) & 0xffu);
            if (cnt) {
                size_t pos = (size_t)(gbase + before + (int)((incl >> (8 * g)) & 0xffu) - cnt);
                const int p = pb + (g * TMA_CONSUMERS + tid) * 4;
                const int row = p / S, col = p - row * S;             // 4 | S: the 4 pixels share the row
                const float yv = (float)__double2ll_rz(__dadd_rn(__dmul_rn(ry, (double)row), y0));
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if ((mbits[g] >> j) & 1u) {
                        if (pos < cap) {
                            float4 P = __ldg(tab + code[g][j]);
                            cb[pos] = (float)__double2ll_rz(__dadd_rn(__dmul_rn(rx, (double)(col + j)), x0));
                            cb[cap + pos] = yv;
                            cb[2 * cap + pos] = P.x;
                            cb[3 * cap + pos] = P.y;
                            cb[4 * cap + pos] = P.z;
                        }
                        pos++;
                    }
                }
            }
            gbase += total;
        }
        running = gbase;
    }
    if (rank == parts - 1 && tid == 0) a.counts[b] = running;
}

template <int DT>
static int launch_tma(zp_ctx* ctx, const DecodeArgs& a, int parts, int part_px, cudaStream_t st) {
    const size_t smem = (size_t)TMA_SLOTS * TMA_SLOT_BYTES;
    static bool smem_set_dev[ZP_MAX_DEVICES] = {};
    bool& smem_set = smem_set_dev[ctx->device % ZP_MAX_DEVICES];
    if (!smem_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_tma_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.B * parts));
    cfg.blockDim = dim3(TMA_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = parts;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ZP_CUDA(ctx, cudaLaunchKernelEx(&cfg, zp_decode_tma_kernel<DT>, a, parts, part_px));
    ctx->launches++;
    return 0;
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
            const int obj = zp_obj_slot(a.obj_ids, a.obj_default, b);
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
static int launch_cluster(zp_ctx* ctx, const DecodeArgs& a, int csize, bool use_cluster, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.B * csize));
    cfg.blockDim = dim3(DEC_THREADS);
    const int smem = 5 * DEC_THREADS * Px<DT>::N * (int)sizeof(float);
    static bool attr_set_dev[ZP_MAX_DEVICES] = {};          // cudaFuncSetAttribute is per device
    bool& attr_set = attr_set_dev[ctx->device % ZP_MAX_DEVICES];
    if (!attr_set) {
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_cluster_kernel<DT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_cluster_kernel<DT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_cluster_kernel<DT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ZP_CUDA(ctx, cudaFuncSetAttribute(zp_decode_cluster_kernel<DT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = use_cluster ? csize : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const bool full16 = a.nb == 16 && a.ext_mask == nullptr;
    if (use_cluster && full16) ZP_CUDA(ctx, cudaLaunchKernelEx(&cfg, zp_decode_cluster_kernel<DT, true, true>, a, csize));
    else if (use_cluster) ZP_CUDA(ctx, cudaLaunchKernelEx(&cfg, zp_decode_cluster_kernel<DT, true, false>, a, csize));
    else if (full16) ZP_CUDA(ctx, cudaLaunchKernelEx(&cfg, zp_decode_cluster_kernel<DT, false, true>, a, csize));
    else ZP_CUDA(ctx, cudaLaunchKernelEx(&cfg, zp_decode_cluster_kernel<DT, false, false>, a, csize));
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
    a.dbg = (unsigned long long*)ctx->dbg_buf;
    const int ppt = dtype == ZP_DTYPE_F32 ? 4 : 8;
    const int esz = dtype == ZP_DTYPE_F32 ? 4 : 2;
    const int N = S * S;
    const int per_cta = DEC_THREADS * ppt;
    const int ctas = (N + per_cta - 1) / per_cta;
    bool vec_ok = a.sw == 1 && S % ppt == 0 && S <= 1024 && a.sh % ppt == 0 && a.sc % ppt == 0 && a.sb % ppt == 0 &&
                  ((uintptr_t)logits % 16) == 0 && ctas <= (ctx->force_decode_path == 1 ? DEC_MAX_CLUSTER : DEC_MAX_RUNS) &&
                  (codes == nullptr || ((uintptr_t)codes % 16) == 0);
    // streaming (TMA) path: contiguous planes, rows that hold whole 4-pixel groups, 16-byte aligned plane segments
    const bool tma_ok = a.sw == 1 && a.sh == S && S % 4 == 0 && (N * esz) % 16 == 0 && ((size_t)a.sc * esz) % 16 == 0 &&
                        ((size_t)a.sb * esz) % 16 == 0 && ((uintptr_t)logits % 16) == 0 &&
                        (codes == nullptr || ((uintptr_t)codes % 8) == 0) && (ext_mask == nullptr || ((uintptr_t)ext_mask % 4) == 0);
    if (tma_ok && ctx->force_decode_path == 3) {
        // one wave of 2 CTAs per SM when the batch is small: split crops until the grid fills it
        int parts = 1;
        const int slots = 2 * ctx->sm_count;
        const int blk = TMA_SLOT_BYTES / esz;
        const int max_parts = (N + blk - 1) / blk;
        while (parts * 2 <= TMA_MAX_PARTS && parts * 2 <= max_parts && B * parts * 2 <= slots) parts *= 2;
        int part_px = (N + parts - 1) / parts;
        part_px = ((part_px + blk - 1) / blk) * blk;
        if (dtype == ZP_DTYPE_F32) return launch_tma<ZP_DTYPE_F32>(ctx, a, parts, part_px, st);
        return launch_tma<ZP_DTYPE_BF16>(ctx, a, parts, part_px, st);
    }
    // two-kernel path: rows of whole PPT-pixel groups, 16-byte aligned vectors; the emit kernel's base pre-count reads
    // <= runs * 16 * PPT words per CTA, so cap the runs per crop
    const bool split_ok = a.sw == 1 && S % ppt == 0 && S <= 1024 && a.sh % ppt == 0 && a.sc % ppt == 0 && a.sb % ppt == 0 &&
                          ((uintptr_t)logits % 16) == 0 && ctas <= 64 && (codes == nullptr || ((uintptr_t)codes % 16) == 0);
    if (split_ok && (ctx->force_decode_path == 4 || (ctx->force_decode_path == 0 && !vec_ok))) {
        if (dtype == ZP_DTYPE_F32) return launch_split<ZP_DTYPE_F32>(ctx, a, st);
        return launch_split<ZP_DTYPE_BF16>(ctx, a, st);
    }
    if (vec_ok && ctx->force_decode_path == 0) {
        if (dtype == ZP_DTYPE_F32) return launch_stream<ZP_DTYPE_F32>(ctx, a, ctas, st);
        return launch_stream<ZP_DTYPE_BF16>(ctx, a, ctas, st);
    }
    if (vec_ok && ctx->force_decode_path != 2) {
        const bool use_cluster = ctx->force_decode_path == 1;
        int csize = ctas;
        if (use_cluster) { csize = 1; while (csize < ctas) csize <<= 1; }
        if (dtype == ZP_DTYPE_F32) return launch_cluster<ZP_DTYPE_F32>(ctx, a, csize, use_cluster, st);
        return launch_cluster<ZP_DTYPE_BF16>(ctx, a, csize, use_cluster, st);
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
