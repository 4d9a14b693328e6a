// Masked flash attention with an ONLINE row maximum: the kernel for configurations whose logits have no a-priori bound.
//
// Replaces fit/model/modules.py:176-204 of the reference:
//     attn_mask[b,i,j] = (mask[b,i] == mask[b,j])                (segment-id equality, NOT a pad mask)
//     o = softmax(q k^T / sqrt(dh) + attn_mask) v ;  o *= (mask != 0)
//
// attention_tm.cuh / attention_ws.cuh (the FiTv2 production kernels) replace the running maximum by a constant bound that
// only holds for affine-free QK-LayerNorm.  FiTv1-style models (q_norm = k_norm = None, configs/fit/config_fit_xl.yaml) and the
// weighted norms (w_layernorm / rmsnorm, fit/model/norms.py:35-50) have unbounded logits, so this kernel keeps the classic
// flash-attention recurrence per query row:
//     m' = max(m, rowmax(S_t));  alpha = 2^(m - m');  l = l * alpha + rowsum(P_t);  O = O * alpha + P_t V_t
// with P_t = 2^(S_t - m').  O is accumulated in REGISTERS (each key tile's P V product is a fresh TMEM accumulation that is
// read back and folded in), so no TMEM rescale pass is needed.
//
// Persistent kernel: CTAs loop over work items (sample, head, 128-query tile); keys are consumed in tiles of 128.  Both
// contractions run on tcgen05:
//     S = Q K^T   : M=128 queries, N=128 keys, K=DHP (head_dim padded to a multiple of 16)
//     P V         : M=128 queries, N=DHP,      K=128 keys   (V arrives transposed, keys contiguous)
// Q/K/V^T tiles are loaded by TMA (3-D tensor maps over (dim, token, sample*head); out-of-bounds rows/columns
// are zero-filled, which pads head_dim 72 -> 80 and the sequence tail for free) straight into the canonical
// K-major swizzled panels: a 64-element SWIZZLE_128B panel plus a 16/32-element SWIZZLE_32B/64B tail panel for
// head_dim 72/96.  Every buffer is refilled as soon as its last reader has retired.
// 256 threads: two threads per query row (warps w and w+4 share a TMEM lane quarter), each owning 64 of the 128
// key columns of a tile; the two partial row maxima / sums meet in shared memory.
// The debug taps (raw S / P V tiles of the first work item) of fitv2_debug_attention live here.
#pragma once
#include "common.cuh"
#include "tc2sm.cuh"

namespace fitv2 {

template <int DH> struct AttnGenCfg {
    static constexpr int kDHP = (DH + 15) / 16 * 16;          // 72 -> 80, 96 -> 96
    static constexpr int kTail = kDHP - 64;                    // elements in the tail panel (16 or 32)
    static constexpr int kTailBytes = kTail * 2;               // 32 or 64
    static constexpr int kQMain = 128 * 128, kQTail = 128 * kTailBytes;
    static constexpr int kOffQ = 0;
    static constexpr int kOffQT = kOffQ + kQMain;
    static constexpr int kOffK = kOffQT + ((kQTail + 1023) / 1024) * 1024;
    static constexpr int kOffKT = kOffK + kQMain;
    static constexpr int kOffV = kOffKT + ((kQTail + 1023) / 1024) * 1024;
    static constexpr int kVPanel = ((kDHP * 128 + 1023) / 1024) * 1024;   // one 64-key panel of V^T
    static constexpr int kOffP = kOffV + 2 * kVPanel;
    static constexpr int kPPanel = 128 * 128;
    static constexpr int kOffSeg = kOffP + 2 * kPPanel;        // 128 key segment ids
    static constexpr int kOffSum = kOffSeg + 128 * 4;          // 256 partial row sums
    static constexpr int kOffMax = kOffSum + 256 * 4;         // 256 partial row maxima
    static constexpr int kOffBar = kOffMax + 256 * 4;
    static constexpr int kSmemBytes = kOffBar + 64 + 1024;     // + alignment slack
    static constexpr uint32_t kQKBytes = kQMain + kQTail;      // bytes of one Q or K tile (TMA writes full boxes)
    static constexpr uint32_t kVBytes = 2 * kDHP * 128;
    static constexpr int kThreads = 256;
    static_assert(kTail == 16 || kTail == 32, "head_dim must be 72..80 or 88..96 (64 + 16/32 tail)");
};

template <typename OT, int DH>
__global__ void __launch_bounds__(256, 2)
attention_general_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qt,
                 const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_kt,
                 const __grid_constant__ CUtensorMap map_v,
                 const float* __restrict__ seg, const int* __restrict__ seg_uniform,
                 OT* __restrict__ out, int heads, int tokens, int num_items, float scale_log2e,
                 float* __restrict__ dbg_s, float* __restrict__ dbg_o)
{
    using C = AttnGenCfg<DH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* seg_kv = reinterpret_cast<float*>(smem + C::kOffSeg);
    float* l_part = reinterpret_cast<float*>(smem + C::kOffSum);
    float* m_part = reinterpret_cast<float*>(smem + C::kOffMax);
    uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + C::kOffBar);   // S = Q K^T done          (one phase per key tile)
    uint64_t* bar_o = bar_s + 1;                                        // O += P V done
    uint64_t* bar_k = bar_s + 2;                                        // K tile (and Q for tile 0) landed
    uint64_t* bar_v = bar_s + 3;                                        // V^T tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 4);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127;                                          // query row inside the tile == TMEM lane
    const int half = tid >> 7;                                          // which 64 key columns / output half this thread owns
    const int q_tiles = (tokens + 127) / 128, kv_tiles = q_tiles;

    // work item -> (query tile, sample*head); query tiles of one (sample, head) are adjacent so K / V stay in L2
    auto issue_qk0 = [&](int item) {                                    // tid 0 only
        const int bh = item / q_tiles, q0 = (item - bh * q_tiles) * 128;
        mbar_arrive_expect_tx(bar_k, 2 * C::kQKBytes);
        tma_load_3d(&map_q, bar_k, smem + C::kOffQ, 0, q0, bh);
        tma_load_3d(&map_qt, bar_k, smem + C::kOffQT, 64, q0, bh);
        tma_load_3d(&map_k, bar_k, smem + C::kOffK, 0, 0, bh);
        tma_load_3d(&map_kt, bar_k, smem + C::kOffKT, 64, 0, bh);
    };
    auto issue_k = [&](int bh, int kv0) {
        mbar_arrive_expect_tx(bar_k, C::kQKBytes);
        tma_load_3d(&map_k, bar_k, smem + C::kOffK, 0, kv0, bh);
        tma_load_3d(&map_kt, bar_k, smem + C::kOffKT, 64, kv0, bh);
    };
    auto issue_v = [&](int bh, int kv0) {
        mbar_arrive_expect_tx(bar_v, C::kVBytes);
        tma_load_3d(&map_v, bar_v, smem + C::kOffV, kv0, 0, bh);
        tma_load_3d(&map_v, bar_v, smem + C::kOffV + C::kVPanel, kv0 + 64, 0, bh);
    };

    if (tid == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
        mbar_init(bar_s, 1); mbar_init(bar_o, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
        mbar_fence_init();
        pdl_wait();                                                     // Q / K / V^T come from the QKV GEMM in front
        if ((int)blockIdx.x < num_items) { issue_qk0(blockIdx.x); issue_v(blockIdx.x / q_tiles, 0); }
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    pdl_wait();
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_s = tmem_base + (uint32_t((warp & 3) * 32) << 16) + half * 64;   // S columns [0,128): this thread's 64
    const uint32_t t_o = tmem_base + (uint32_t((warp & 3) * 32) << 16) + 128;         // O columns [128,128+DHP)
    const uint32_t smem_p = smem_u32(smem + C::kOffP);
    constexpr uint32_t idesc_s = umma_idesc(Op16<OT>::kUmmaFormat, 128, 128);
    constexpr uint32_t idesc_o = umma_idesc(Op16<OT>::kUmmaFormat, 128, C::kDHP);
    uint32_t ph = 0;                                                    // parity of the running key-tile counter

    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int bh = item / q_tiles, q0 = (item - bh * q_tiles) * 128;
        const int sample = bh / heads, head = bh - sample * heads;
        const int next_item = item + gridDim.x;
        const float* segb = seg + (size_t)sample * tokens;
        const bool uniform = seg_uniform[sample] == tokens;            // key lengths < tokens (padded samples) take the compare path here
        const int qi = q0 + row;
        const bool q_ok = qi < tokens;
        const float my_seg = q_ok ? segb[qi] : 0.f;
        constexpr int OH = C::kDHP / 2;                                // output columns per thread: 40 or 48
        float l_run = 0.f, m_run = -INFINITY;                          // m_run in log2 units (already scaled)
        float o_acc[OH];
#pragma unroll
        for (int j = 0; j < OH; ++j) o_acc[j] = 0.f;

        for (int t = 0; t < kv_tiles; ++t, ph ^= 1) {
            const int kv0 = t * 128;
            if (tid < 128) seg_kv[tid] = (kv0 + tid < tokens) ? segb[kv0 + tid] : 0.f;

            // ---- S = Q K^T ----
            if (tid == 0) {
                mbar_wait(bar_k, ph);                                  // K(t) (and Q) landed
                tc_fence_after();
                const uint64_t dq = umma_desc_kmajor(smem_u32(smem + C::kOffQ), 128);
                const uint64_t dk = umma_desc_kmajor(smem_u32(smem + C::kOffK), 128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_ss(tmem_base, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
                const uint64_t dqt = umma_desc_kmajor(smem_u32(smem + C::kOffQT), C::kTailBytes);
                const uint64_t dkt = umma_desc_kmajor(smem_u32(smem + C::kOffKT), C::kTailBytes);
#pragma unroll
                for (int kk = 0; kk < C::kTail / 16; ++kk) umma_ss(tmem_base, dqt + 2 * kk, dkt + 2 * kk, idesc_s, 1);
                umma_commit(bar_s);
            }
            __syncthreads();                                           // seg_kv visible; previous item's epilogue finished
            mbar_wait(bar_s, ph);
            tc_fence_after();
            if (tid == 0) {                                            // K (and, after the last tile, Q) buffers are free
                if (t + 1 < kv_tiles) issue_k(bh, kv0 + 128);
                else if (next_item < num_items) issue_qk0(next_item);
            }

            const int kv_valid = min(128, tokens - kv0);
            const bool dense = uniform && kv_valid == 128;             // CTA-uniform: no per-element masking needed
            // ---- pass 1: masked row maximum of this thread's 64 columns ----
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(t_s + c * 32, v);
                tmem_ld_wait();
                const int colbase = half * 64 + c * 32;
                if (dbg_s != nullptr && t == 0 && item == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) dbg_s[row * 128 + colbase + j] = __uint_as_float(v[j]);   // raw S tile (debug)
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = colbase + j;
                    const bool ok = dense || (col < kv_valid && (uniform || seg_kv[col] == my_seg));
                    if (ok) mx = fmaxf(mx, __uint_as_float(v[j]) * scale_log2e);
                }
            }
            m_part[tid] = mx;
            __syncthreads();
            const float m_new = fmaxf(m_run, fmaxf(m_part[row], m_part[row + 128]));
            const float m_safe = m_new == -INFINITY ? 0.f : m_new;     // nothing visible yet: every p below is masked to 0
            const float alpha = fast_exp2(m_run - m_safe);             // m_run = -inf -> 0
            m_run = m_new;

            // ---- pass 2: p = exp2(s*c - m), row sum, P -> smem (A operand of P V) ----
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(t_s + c * 32, v);
                tmem_ld_wait();
                const int colbase = half * 64 + c * 32;
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = colbase + 2 * j;
                    const bool ok0 = dense || (col < kv_valid && (uniform || seg_kv[col] == my_seg));
                    const bool ok1 = dense || (col + 1 < kv_valid && (uniform || seg_kv[col + 1] == my_seg));
                    const float p0 = ok0 ? fast_exp2(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -m_safe)) : 0.f;
                    const float p1 = ok1 ? fast_exp2(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -m_safe)) : 0.f;
                    packed[j] = Op16<OT>::pack(p0, p1);
                    lsum += p0 + p1;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {                          // 4 chunks of 8 keys per 32 columns; panel = this thread's half
                    const int chunk = c * 4 + g;                        // 0..7 inside the 64-key panel
                    sts128(smem_p + half * C::kPPanel + swz_offset<128>(row, chunk),
                           make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]));
                }
            }
            l_run = l_run * alpha + lsum;
            tc_fence_before();
            fence_proxy_async_smem();
            __syncthreads();

            // ---- O_t = P V (fresh accumulation), folded into the register accumulator ----
            if (tid == 0) {
                mbar_wait(bar_v, ph);                                  // V(t) landed
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t dp = umma_desc_kmajor(smem_u32(smem + C::kOffP + (kk >> 2) * C::kPPanel), 128) + 2 * (kk & 3);
                    const uint64_t dv = umma_desc_kmajor(smem_u32(smem + C::kOffV + (kk >> 2) * C::kVPanel), 128) + 2 * (kk & 3);
                    umma_ss(tmem_base + 128, dp, dv, idesc_o, kk != 0);
                }
                umma_commit(bar_o);
            }
            mbar_wait(bar_o, ph);                                      // P and V buffers are free, O_t complete
            tc_fence_after();
            if (tid == 0) {
                if (t + 1 < kv_tiles) issue_v(bh, kv0 + 128);
                else if (next_item < num_items) issue_v(next_item / q_tiles, 0);
            }
            float o[OH];
            tmem_ld32(t_o + half * OH, reinterpret_cast<uint32_t*>(o));
            if constexpr (OH == 40) tmem_ld8(t_o + half * OH + 32, reinterpret_cast<uint32_t*>(o) + 32);
            else tmem_ld16(t_o + half * OH + 32, reinterpret_cast<uint32_t*>(o) + 32);
            tmem_ld_wait();
            if (dbg_o != nullptr && item == 0 && t == 0) {
#pragma unroll
                for (int j = 0; j < OH; ++j) dbg_o[row * C::kDHP + half * OH + j] = o[j];                // raw P V tile (debug)
            }
#pragma unroll
            for (int j = 0; j < OH; ++j) o_acc[j] = fmaf(o_acc[j], alpha, o[j]);
            tc_fence_before();                                         // the next tile's P V overwrites the O columns
        }

        // ---- O / l, zero padded queries (mask != 0), write (M, heads*DH) rows for the proj GEMM ----
        l_part[tid] = l_run;
        __syncthreads();
        const float l_tot = l_part[row] + l_part[row + 128];
        if (q_ok) {
            const float inv = (my_seg != 0.f && l_tot > 0.f) ? 1.0f / l_tot : 0.f;
            OT* dst = out + ((size_t)sample * tokens + qi) * (heads * DH) + head * DH + half * OH;
#pragma unroll
            for (int c = 0; c < OH / 8; ++c) {
                if (half * OH + c * 8 < DH) {                          // skip the zero-pad columns 72..79
                    uint32_t pk[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) pk[p] = Op16<OT>::pack(o_acc[c * 8 + 2 * p] * inv, o_acc[c * 8 + 2 * p + 1] * inv);
                    *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
        // l_part / seg_kv are rewritten only after the next item's first __syncthreads
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace fitv2
