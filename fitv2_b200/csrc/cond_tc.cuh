// Conditioning linears on the tensor pipe (tcgen05 kind::tf32): global adaLN, final adaLN, adaLN-LoRA down and the
// batched adaLN-LoRA up (fit/model/fit_model.py:218-219, fit/model/modules.py:259-264,287-293).
//
//   out[z][r, n] = sum_k A[z][r, k] * W[z][n, k] + bias[z][n] (+ add[r, n])        r < rows (64 with CFG), fp32 in / out
//
// These are weight-bandwidth-bound (376 MB of fp32 weights per NFE at XL/2 against 12 GFLOP), but the fp32-FMA SIMT kernel
// of pointwise.cuh is bound by the FMA pipe (0.55 ms of the 0.65 ms conditioning phase).  Here the weights stream through
// TMA straight into the UMMA operand layout (fp32 in shared memory IS the tf32 operand: the tensor pipe reads the upper
// 19 bits) and precision is kept on the activation side with a hi / lo split:
//   a = hi + lo,  hi = tf32(a) (round to nearest),  lo = a - hi (exact in fp32, |lo| <= 2^-11 |a|)
// The two parts are STACKED along M: rows 0..63 of the 128-row A tile hold hi, rows 64..127 hold lo of the same 64 logical
// rows, so ONE M = 128 UMMA computes both products and the epilogue adds TMEM lane r + 64 to lane r.  The weights are
// used at tf32 precision (the host packer rounds them to nearest once; unrounded fp32 weights are truncated by the tensor
// pipe), i.e. 11 significant bits (2^-11 relative rounding) against the 8 of the bf16 block GEMMs that consume the modulation.
//
//   warp 0 : TMA producer (A split tile 128 x 32 fp32 + weight tile BN x 32 fp32 per stage, 128B swizzle)
//   warp 1 : TMEM allocator + tcgen05.mma issuer (4 UMMAs of K = 8 per stage), double-buffered accumulators
//   warps 2-5 : epilogue.  lo lanes -> shared staging tile, hi lanes add, then a coalesced write-out of the 64 x BN tile
//               with bias / add; optionally the hi / lo split of the result for the next linear (adaLN-LoRA down -> up).
// Persistent over (batch, m-tile, n-tile) items; up to three weight matrices that share A are served by one launch.
#pragma once
#include "common.cuh"
#include "attention_ws.cuh"   // named_bar_sync

namespace fitv2 {

constexpr int kCondSegs = 3;
constexpr int kCondRows = 64;                 // logical rows per tile (hi + lo = 128 UMMA rows)
constexpr int kCondBK = 32;                   // fp32 elements per 128-byte swizzle row

struct CondTc {
    int nseg, rows, K, batches;
    int a_batch_cols;                         // column offset of batch z inside the A split tensor (K of one batch)
    int nt_prefix[kCondSegs + 1];             // prefix sums of N_s / BN
    int w_batch_rows[kCondSegs];              // row offset of batch z inside weight tensor s (N_s)
    const float* bias[kCondSegs]; int bias_batch_stride[kCondSegs];
    const float* add[kCondSegs];              // (rows, N_s) or null, shared by all batches
    float* out[kCondSegs]; size_t out_batch_stride[kCondSegs]; int ldo[kCondSegs];
    float* out_split[kCondSegs]; int ld_split[kCondSegs];   // optional [m_tiles][128][ld_split]: rows 0..63 hi, 64..127 lo
};

// D[tmem] (+)= A[smem] * B[smem], kind::tf32 (fp32 storage, 10-bit mantissa operands, fp32 accumulate)
__device__ __forceinline__ void umma_ss_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

template <int BN> struct CondCfg {
    static constexpr int kABytes = 128 * kCondBK * 4;                  // 16 KB
    static constexpr int kBBytes = BN * kCondBK * 4;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kPitch = BN + 4;                              // staging row pitch (floats): conflict-free 128-bit rows
    static constexpr int kEpiBytes = kCondRows * kPitch * 4;
    static constexpr int kStagesRaw = (200 * 1024 - kEpiBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 + 1024;
    static constexpr int kAccStride = 128;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 128, "tile width");
    static_assert(kBBytes % 1024 == 0, "weight stage must keep the 1024-byte swizzle-atom alignment");
    static_assert(kStages >= 3, "pipeline too shallow");
};

struct CondItem { int z, m_tile, seg, n0; };

__device__ __forceinline__ CondItem cond_decode(const CondTc& p, int tile, int m_tiles, int bn) {
    const int nt_total = p.nt_prefix[p.nseg];
    const int per_batch = m_tiles * nt_total;
    CondItem it;
    it.z = tile / per_batch;
    const int rem = tile - it.z * per_batch;
    it.m_tile = rem / nt_total;
    const int nt = rem - it.m_tile * nt_total;
    it.seg = (p.nseg > 1 && nt >= p.nt_prefix[1]) ? ((p.nseg > 2 && nt >= p.nt_prefix[2]) ? 2 : 1) : 0;
    it.n0 = (nt - p.nt_prefix[it.seg]) * bn;
    return it;
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
cond_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w0,
               const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const CondTc p)
{
    using Cfg = CondCfg<BN>;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
    float* stage_s = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes + Cfg::kEpiBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.rows + kCondRows - 1) / kCondRows;
    const int num_tiles = p.batches * m_tiles * p.nt_prefix[p.nseg];
    const int num_kb = (p.K + kCondBK - 1) / kCondBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w0);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const CondItem it = cond_decode(p, tile, m_tiles, BN);
            const CUtensorMap* mw = it.seg == 0 ? &map_w0 : (it.seg == 1 ? &map_w1 : &map_w2);
            const int w_row = it.z * p.w_batch_rows[it.seg] + it.n0;
            const int a_col = it.z * p.a_batch_cols;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_2d(&map_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, a_col + kb * kCondBK, it.m_tile * 128);
                    tma_load_2d(mw, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kCondBK, w_row);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        constexpr uint32_t idesc = umma_idesc(2u /* TF32 */, 128, BN);
        const uint64_t da0 = umma_desc_kmajor(smem_u32(smem_a), 128);
        const uint64_t db0 = umma_desc_kmajor(smem_u32(smem_b), 128);
        int stage = 0; uint32_t phase = 0; int n_it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n_it) {
            const int acc = n_it & 1;
            const uint32_t acc_phase = (n_it >> 1) & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = da0 + (uint64_t)(stage * (Cfg::kABytes >> 4));
                    const uint64_t db = db0 + (uint64_t)(stage * (Cfg::kBBytes >> 4));
#pragma unroll
                    for (int kk = 0; kk < kCondBK / 8; ++kk)                  // +32 bytes (>>4 = 2) per K = 8 step inside the 128B atom
                        umma_ss_tf32(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0);
                    umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue (4 warps) ------------------------------
        const int quarter = warp & 3;                                   // TMEM lane quarter this warp may access
        const bool is_lo = quarter >= 2;                                // lanes 64..127 hold the lo products
        const int r_loc = (quarter & 1) * 32 + lane;                    // logical row inside the 64-row tile
        const int et = threadIdx.x - 64;                                // 0..127
        float* srow = stage_s + r_loc * Cfg::kPitch;
        int n_it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n_it) {
            const CondItem it = cond_decode(p, tile, m_tiles, BN);
            const int acc = n_it & 1;
            const uint32_t acc_phase = (n_it >> 1) & 1;
            const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (uint32_t(quarter * 32) << 16);
            // the `add` operand (global adaLN, L2-resident) does not depend on the accumulator: request the whole tile's share
            // before waiting for the MMAs (the write-out loop below was bound by these loads: 8.7 us per tile)
            constexpr int CPR = BN / 4;                                 // float4 chunks per row
            constexpr int ITERS = kCondRows * CPR / 128;
            static_assert(kCondRows * CPR % 128 == 0, "write-out mapping");
            const int seg = it.seg;
            const int n_seg = (p.nt_prefix[seg + 1] - p.nt_prefix[seg]) * BN;
            const float* add = p.add[seg] ? p.add[seg] + it.n0 : nullptr;
            float4 addv[ITERS];
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                const int idx = et + i * 128, r = idx / CPR, c = (idx - r * CPR) * 4;
                const int rg = it.m_tile * kCondRows + r;
                addv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (add && rg < p.rows) addv[i] = *reinterpret_cast<const float4*>(add + (size_t)rg * n_seg + c);
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (is_lo) {
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(srow + c0 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
                named_bar_sync(1, 128);                                 // lo parts staged
            } else {
                named_bar_sync(1, 128);
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 s = *reinterpret_cast<const float4*>(srow + c0 + 4 * j);
                        s.x += __uint_as_float(v[4 * j]); s.y += __uint_as_float(v[4 * j + 1]);
                        s.z += __uint_as_float(v[4 * j + 2]); s.w += __uint_as_float(v[4 * j + 3]);
                        *reinterpret_cast<float4*>(srow + c0 + 4 * j) = s;
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
            }
            named_bar_sync(2, 128);                                     // hi + lo sums staged
            // coalesced write-out of the 64 x BN tile
            const float* bias = p.bias[seg] ? p.bias[seg] + (size_t)it.z * p.bias_batch_stride[seg] + it.n0 : nullptr;
            float* out = p.out[seg] + (size_t)it.z * p.out_batch_stride[seg] + it.n0;
            float* split = p.out_split[seg] ? p.out_split[seg] + (size_t)it.m_tile * 128 * p.ld_split[seg] + it.n0 : nullptr;
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                const int idx = et + i * 128, r = idx / CPR, c = (idx - r * CPR) * 4;
                const int rg = it.m_tile * kCondRows + r;
                if (rg >= p.rows) continue;
                float4 v = *reinterpret_cast<const float4*>(stage_s + r * Cfg::kPitch + c);
                if (bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c)); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
                v.x += addv[i].x; v.y += addv[i].y; v.z += addv[i].z; v.w += addv[i].w;
                *reinterpret_cast<float4*>(out + (size_t)rg * p.ldo[seg] + c) = v;
                if (split) {
                    float4 hi = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                    float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    *reinterpret_cast<float4*>(split + (size_t)r * p.ld_split[seg] + c) = hi;
                    *reinterpret_cast<float4*>(split + (size_t)(r + 64) * p.ld_split[seg] + c) = lo;
                }
            }
            named_bar_sync(3, 128);                                     // staging tile free for the next item
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace fitv2
