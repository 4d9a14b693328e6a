// Attention variant with P kept in TENSOR MEMORY: the production kernel up to 256 tokens at head_dim 72 (FITV2_ATTN=ws / tm
// overrides the choice; beyond 256 tokens attention_ws.cuh is faster, see profiles/README.md).
//
// Same math, operand layouts and roles as attention_ws.cuh.  Differences:
//   * key tiles are consumed in 64-key SUB-tiles; every stream owns two 64-column S buffers in TMEM (ping-pong).
//     The softmax warps read S(k) out of buffer k % 2, and write the 16-bit P(k) back INTO the first 32 columns of the
//     same buffer with tcgen05.st (each thread over the first 16 of its OWN 32 S columns, so no cross-warp hazard);
//     O += P V is a tcgen05.mma with the A operand taken from TMEM.  No shared-memory
//     round trip for P (st.shared + fence.proxy.async + 256 KB of the 690 KB of shared-memory traffic per work item).
//   * S(k+2) is issued right behind P V(k) by the same thread, so the in-order tensor pipe orders the reuse of the
//     buffer and S always runs two sub-tiles ahead of the softmax: the softmax warps never wait for S.
//   * Q is double-buffered per stream (the P buffers freed 64 KB of shared memory).
// TMEM per stream: S0 [0,64) S1 [64,128) O [128,128+DHP); stream b starts at column 256.
#pragma once
#include "attention_ws.cuh"

namespace fitv2 {

// TPR = softmax threads per query row: 2 (eight warps per stream, 32 S columns per thread and sub-tile) or 1 (four warps
// per stream, one thread owns the whole row: half as many barrier / TMEM round trips per exponential, no row-sum exchange).
template <int DH, int TPR = 2> struct AttnTmCfg {
    using W = AttnWsCfg<DH>;
    static constexpr int kDHP = W::kDHP, kTail = W::kTail, kTailBytes = W::kTailBytes, kQMain = W::kQMain;
    static constexpr int kQKTile = W::kQKTile, kVPanel = W::kVPanel, kVTile = W::kVTile;
    static constexpr int kKStages = 2, kVStages = 2;      // (three stages + the double-buffered Q exceed 227 KB at head_dim 72)
    static constexpr int kOffQ = 0;                                      // [stream][buffer]
    static constexpr int kOffK = kOffQ + 4 * kQKTile;
    static constexpr int kOffV = kOffK + kKStages * kQKTile;
    static constexpr int kOffStage = kOffV + kVStages * kVTile;          // output staging tile per stream
    static constexpr int kStageTile = ((128 * DH * 2 + 1023) / 1024) * 1024;
    static constexpr int kOffSum = kOffStage + 2 * kStageTile;
    static constexpr int kOffBar = kOffSum + 2 * 256 * 4;
    static constexpr int kNumBars = 40;
    static constexpr int kOffSeg = kOffBar + ((kNumBars * 8 + 16 + 127) / 128) * 128;
    static constexpr uint32_t kQKBytes = W::kQKBytes, kVBytes = W::kVBytes;
    static constexpr int kSoftWarps = 4 * TPR;            // per stream
    static constexpr int kThreads = 32 * (3 + 2 * kSoftWarps);
    static constexpr int smem_bytes(int tokens) { return kOffSeg + 2 * ((tokens + 127) / 128 * 128) * 4 + 1024; }
};

// D[tmem] (+)= A[tmem] * B[smem]: A rows = TMEM lanes, two consecutive K elements per 32-bit column
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    tmem_st16(taddr, v);
    tmem_st16(taddr + 16, v + 16);
}

template <typename OT, int DH, int TPR = 2>
__global__ void __launch_bounds__((AttnTmCfg<DH, TPR>::kThreads), 1)
attention_tm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qt,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_kt,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                    const float* __restrict__ seg, const int* __restrict__ seg_uniform,
                    int heads, int tokens, int num_items, float scale_log2e, float bound_log2e, int early)
{
    using C = AttnTmCfg<DH, TPR>;
    constexpr int SW = C::kSoftWarps, SOFT_THREADS = 32 * SW, COLS = 64 / TPR;   // per stream; S columns per thread and sub-tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;                    // [stream][buffer]  Q landed
    uint64_t* q_empty = bars + 4;               // [stream][buffer]  every S MMA of the item has retired
    uint64_t* k_full = bars + 8;                // [kKStages]
    uint64_t* k_empty = k_full + C::kKStages;   //                   both streams have released the stage (2 arrivals)
    uint64_t* v_full = k_empty + C::kKStages;
    uint64_t* v_empty = v_full + C::kVStages;
    uint64_t* s_full = v_empty + C::kVStages;   // [stream][buffer]  S sub-tile accumulated in TMEM
    uint64_t* p_full = s_full + 4;              // [stream][buffer]  P sub-tile written back to TMEM (8 warp arrivals)
    uint64_t* o_full = p_full + 4;              // [stream]          last P V of the item retired
    uint64_t* o_free = o_full + 2;              // [stream]          softmax warps have read O (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tiles = (tokens + 127) / 128, kv_tiles = q_tiles, sub_tiles = 2 * kv_tiles;
    const int q_pairs = (q_tiles + 1) / 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_qt); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_kt);
        tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
        for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&p_full[i], SW); }
        for (int i = 0; i < 2; ++i) { mbar_init(&o_full[i], 1); mbar_init(&o_free[i], SW); }
        for (int i = 0; i < C::kKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }
        for (int i = 0; i < C::kVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ------------------------------------- TMA producer -------------------------------------
        uint32_t ks = 0, kph = 0, vs = 0, vph = 0, n_q[2] = {0, 0};
        auto load_k = [&](int bh, int t) {
            mbar_wait(&k_empty[ks], kph ^ 1);
            if (elect_one()) {
                uint8_t* dst = smem + C::kOffK + ks * C::kQKTile;
                mbar_arrive_expect_tx(&k_full[ks], C::kQKBytes);
                tma_load_3d(&map_k, &k_full[ks], dst, 0, t * 128, bh);
                tma_load_3d(&map_kt, &k_full[ks], dst + C::kQMain, 64, t * 128, bh);
            }
            __syncwarp();
            if (++ks == C::kKStages) { ks = 0; kph ^= 1; }
        };
        auto load_v = [&](int bh, int t) {
            mbar_wait(&v_empty[vs], vph ^ 1);
            if (elect_one()) {
                uint8_t* dst = smem + C::kOffV + vs * C::kVTile;
                mbar_arrive_expect_tx(&v_full[vs], C::kVBytes);
                tma_load_3d(&map_v, &v_full[vs], dst, t * 128, 0, bh);
                tma_load_3d(&map_v, &v_full[vs], dst + C::kVPanel, t * 128 + 64, 0, bh);
            }
            __syncwarp();
            if (++vs == C::kVStages) { vs = 0; vph ^= 1; }
        };
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int bh = item / q_pairs, qp = item - bh * q_pairs;
            const int nstreams = (2 * qp + 1 < q_tiles) ? 2 : 1;
            for (int x = 0; x < nstreams; ++x) {
                const uint32_t qb = n_q[x] & 1;                         // Q buffer of this item; use number n_q >> 1 of that buffer
                if (n_q[x] >= 2) mbar_wait(&q_empty[x * 2 + qb], ((n_q[x] >> 1) - 1) & 1);
                if (elect_one()) {
                    uint8_t* dst = smem + C::kOffQ + (x * 2 + qb) * C::kQKTile;
                    mbar_arrive_expect_tx(&q_full[x * 2 + qb], C::kQKBytes);
                    tma_load_3d(&map_q, &q_full[x * 2 + qb], dst, 0, (2 * qp + x) * 128, bh);
                    tma_load_3d(&map_qt, &q_full[x * 2 + qb], dst + C::kQMain, 64, (2 * qp + x) * 128, bh);
                }
                __syncwarp();
                ++n_q[x];
            }
            for (int t = 0; t < kv_tiles; ++t) { load_k(bh, t); load_v(bh, t); }
        }
    } else if (warp <= 2) {
        // ------------------------------------- tcgen05 issuer of stream x -------------------------------------
        const int x = warp - 1;
        constexpr uint32_t idesc_s = umma_idesc(Op16<OT>::kUmmaFormat, 128, 64);
        constexpr uint32_t idesc_o = umma_idesc(Op16<OT>::kUmmaFormat, 128, C::kDHP);
        const uint32_t sm_q = smem_u32(smem + C::kOffQ) + x * 2 * C::kQKTile, sm_k = smem_u32(smem + C::kOffK), sm_v = smem_u32(smem + C::kOffV);
        const uint32_t t_x = tmem_base + x * 256;
        auto active = [&](int item) { return x == 0 || 2 * (item % q_pairs) + 1 < q_tiles; };
        struct Seq { int item, j; uint32_t n, n_item, stage, phase; };   // j: 64-key sub-tile inside the item; n: active sub-tiles issued
        Seq sq = {(int)blockIdx.x, 0, 0u, 0u, 0u, 0u}, pv = sq;
        auto s_step = [&]() {
            if (sq.item < num_items) {
                const bool act = active(sq.item);
                const int sub = sq.j & 1;
                if (sub == 0) mbar_wait(&k_full[sq.stage], sq.phase);
                if (act) {
                    const uint32_t qb = sq.n_item & 1, b = sq.n & 1;
                    if (sq.j == 0) mbar_wait(&q_full[x * 2 + qb], (sq.n_item >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t dq = umma_desc_kmajor(sm_q + qb * C::kQKTile, 128);
                        const uint64_t dk = umma_desc_kmajor(sm_k + sq.stage * C::kQKTile + sub * 64 * 128, 128);
                        const uint32_t d = t_x + b * 64;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_ss(d, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
                        const uint64_t dqt = umma_desc_kmajor(sm_q + qb * C::kQKTile + C::kQMain, C::kTailBytes);
                        const uint64_t dkt = umma_desc_kmajor(sm_k + sq.stage * C::kQKTile + C::kQMain + sub * 64 * C::kTailBytes, C::kTailBytes);
#pragma unroll
                        for (int kk = 0; kk < C::kTail / 16; ++kk) umma_ss(d, dqt + 2 * kk, dkt + 2 * kk, idesc_s, 1);
                        umma_commit(&s_full[x * 2 + b]);
                        if (sub == 1) umma_commit(&k_empty[sq.stage]);
                        if (sq.j + 1 == sub_tiles) umma_commit(&q_empty[x * 2 + qb]);
                    }
                    __syncwarp();
                    ++sq.n;
                } else if (sub == 1) {
                    if (elect_one()) mbar_arrive(&k_empty[sq.stage]);
                    __syncwarp();
                }
                if (sub == 1) { if (++sq.stage == C::kKStages) { sq.stage = 0; sq.phase ^= 1; } }
                if (++sq.j == sub_tiles) { sq.j = 0; sq.item += gridDim.x; if (act) ++sq.n_item; }
                return;                                                 // one tile per call, also for an idle stream: K and V stages must be
            }                                                           // handed back in the producer's order or the rings deadlock
        };
        auto pv_step = [&]() {
            if (pv.item < num_items) {
                const bool act = active(pv.item);
                const int sub = pv.j & 1;
                if (sub == 0) mbar_wait(&v_full[pv.stage], pv.phase);
                if (act) {
                    const uint32_t b = pv.n & 1;
                    mbar_wait(&p_full[x * 2 + b], (pv.n >> 1) & 1);
                    if (pv.j == 0 && pv.n_item > 0) mbar_wait(&o_free[x], (pv.n_item - 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a = t_x + b * 64, d = t_x + 128;
                        const uint64_t dv = umma_desc_kmajor(sm_v + pv.stage * C::kVTile + sub * C::kVPanel, 128);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)                  // TPR 2: keys 0-31 live in columns [0,16), keys 32-63 in [32,48);
                            umma_ts(d, a + (TPR == 2 ? (kk >> 1) * 32 + (kk & 1) * 8 : kk * 8), dv + 2 * kk, idesc_o, (pv.j | kk) != 0);   // TPR 1: [0,32)
                        if (sub == 1) umma_commit(&v_empty[pv.stage]);
                        if (pv.j + 1 == sub_tiles) umma_commit(&o_full[x]);
                    }
                    __syncwarp();
                    ++pv.n;
                } else if (sub == 1) {
                    if (elect_one()) mbar_arrive(&v_empty[pv.stage]);
                    __syncwarp();
                }
                if (sub == 1) { if (++pv.stage == C::kVStages) { pv.stage = 0; pv.phase ^= 1; } }
                if (++pv.j == sub_tiles) { pv.j = 0; pv.item += gridDim.x; if (act) ++pv.n_item; }
                return;                                                 // one tile per call, also for an idle stream: K and V stages must be
            }                                                           // handed back in the producer's order or the rings deadlock
        };
        s_step(); s_step();                                             // S runs two sub-tiles ahead of P V
        while (pv.item < num_items) { pv_step(); s_step(); }
    } else {
        // ------------------------------------- softmax warp groups -------------------------------------
        const int x = (warp - 3) / SW;
        const int wl = (warp - 3) - x * SW;
        const int half = wl >> 2;                                       // which 32 key columns of the sub-tile / output half (TPR 1: 0)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int tid_wg = wl * 32 + lane;
        const uint32_t t_x = tmem_base + (uint32_t(quarter * 32) << 16) + x * 256;
        const uint32_t stage_sm = smem_u32(smem + C::kOffStage) + x * C::kStageTile;
        float* l_part = reinterpret_cast<float*>(smem + C::kOffSum) + x * 256;
        const int seg_pad = (tokens + 127) / 128 * 128;
        float* seg_s = reinterpret_cast<float*>(smem + C::kOffSeg) + x * seg_pad;
        uint32_t n_s = 0, n_item = 0;
        bool store_pending = false;
        int uni_nx = 1; float seg_nx = 0.f;
        auto fetch_meta = [&](int item) {
            if (item < num_items) {
                const int bh = item / q_pairs, qp = item - bh * q_pairs;
                const int sample = bh / heads, qi = (2 * qp + x) * 128 + row;
                uni_nx = __ldg(seg_uniform + sample);
                seg_nx = qi < tokens ? __ldg(seg + (size_t)sample * tokens + qi) : 0.f;
            }
        };
        fetch_meta(blockIdx.x);
        uint32_t v[COLS];
        // One 64-key sub-tile of the softmax: S(k) out of TMEM buffer n_s % 2, P(k) back into its first 32 columns, hand-over to
        // the issuer.  Returns this thread's partial row sum.
        auto do_sub = [&](int j, int klen, float my_seg) -> float {     // klen: key length of the sample (seg_uniform_kernel), 0 = compare ids
            const bool uniform = klen != 0;
            const int kv0 = j * 64;
            const int kv_valid = min(64, (uniform ? klen : tokens) - kv0);   // may be <= 0 behind the last valid key
            const int mode = (uniform && kv_valid == 64) ? 0 : (uniform ? 1 : 2);
            const uint32_t b = n_s & 1;
            mbar_wait(&s_full[x * 2 + b], (n_s >> 1) & 1);
            tc_fence_after();
            tmem_ld32(t_x + b * 64 + half * 32, v);
            if constexpr (TPR == 1) tmem_ld32(t_x + b * 64 + 32, v + 32);
            tmem_ld_wait();
            uint32_t packed[COLS / 2];
            float lsum = 0.f;
            auto soft32 = [&](auto mode_c) {
                constexpr int kMode = decltype(mode_c)::value;
#pragma unroll
                for (int i = 0; i < COLS / 2; ++i) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), scale_log2e, -bound_log2e));
                    float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2e, -bound_log2e));
                    if constexpr (kMode != 0) {
                        const int col = half * 32 + 2 * i;
                        bool ok0 = col < kv_valid, ok1 = col + 1 < kv_valid;
                        if constexpr (kMode == 2) {
                            ok0 = ok0 && seg_s[kv0 + col] == my_seg;
                            ok1 = ok1 && seg_s[kv0 + col + 1] == my_seg;
                        }
                        p0 = ok0 ? p0 : 0.f;
                        p1 = ok1 ? p1 : 0.f;
                    }
                    packed[i] = Op16<OT>::pack(p0, p1);
                    lsum += p0 + p1;
                }
            };
            if (mode == 0) soft32(std::integral_constant<int, 0>{});
            else if (mode == 1) soft32(std::integral_constant<int, 1>{});
            else soft32(std::integral_constant<int, 2>{});
            if constexpr (TPR == 2) tmem_st16(t_x + b * 64 + half * 32, packed);   // P(k) over the first half of this thread's own S columns
            else tmem_st32(t_x + b * 64, packed);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[x * 2 + b]);
            ++n_s;
            return lsum;
        };
        bool early_done = false;                                        // sub-tile 0 of this item was already processed ...
        float l_early = 0.f;                                            // ... with this partial row sum
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int klen = uni_nx;
            const bool uniform = klen != 0;
            const float my_seg = seg_nx;
            fetch_meta(item + gridDim.x);
            const int bh = item / q_pairs, qp = item - bh * q_pairs;
            const int qt = 2 * qp + x;
            if (qt >= q_tiles) continue;
            const int sample = bh / heads, head = bh - sample * heads;
            if (!uniform) {
                const float* segb = seg + (size_t)sample * tokens;
                for (int i = tid_wg; i < seg_pad; i += SOFT_THREADS) seg_s[i] = i < tokens ? __ldg(segb + i) : 0.f;
                named_bar_sync(1 + x, SOFT_THREADS);
            }
            float l_run = early_done ? l_early : 0.f;
            for (int j = early_done ? 1 : 0; j < sub_tiles; ++j) l_run += do_sub(j, klen, my_seg);
            early_done = false;
            // The last P V of the item has only just been issued: instead of waiting for it (17 % of the warp samples sat in
            // the o_full wait below), the first sub-tile of the NEXT item -- its S was issued two sub-tiles ago -- is done now.
            // The issuer's first P V of that item still waits for o_free, i.e. for the O read of the epilogue below.  Only when
            // the next item needs no segment table (a staged table would be overwritten under the slower warps of this item).
            {
                const int nxt = item + gridDim.x;
                if (early && nxt < num_items && uni_nx != 0) {
                    const int bh_n = nxt / q_pairs;
                    if (2 * (nxt - bh_n * q_pairs) + x < q_tiles) { l_early = do_sub(0, uni_nx, seg_nx); early_done = true; }
                }
            }

            // ---- epilogue ----
            if constexpr (TPR == 2) l_part[half * 128 + row] = l_run;
            if (store_pending) {                                        // the previous TMA store has read the staging tile
                if (tid_wg == 0) tma_store_wait_read();
                store_pending = false;
            }
            mbar_wait(&o_full[x], n_item & 1);
            tc_fence_after();
            constexpr int OH = C::kDHP / TPR;
            float o[OH];
            tmem_ld32(t_x + 128 + half * OH, reinterpret_cast<uint32_t*>(o));
            if constexpr (OH == 40) tmem_ld8(t_x + 128 + half * OH + 32, reinterpret_cast<uint32_t*>(o) + 32);
            else if constexpr (OH == 48) tmem_ld16(t_x + 128 + half * OH + 32, reinterpret_cast<uint32_t*>(o) + 32);
            else {
                static_assert(OH == 40 || OH == 48 || OH == 80, "output columns per thread");
                tmem_ld32(t_x + 128 + 32, reinterpret_cast<uint32_t*>(o) + 32);
                tmem_ld16(t_x + 128 + 64, reinterpret_cast<uint32_t*>(o) + 64);
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[x]);
            named_bar_sync(1 + x, SOFT_THREADS);                        // row sums exchanged; staging tile free
            const float l_tot = TPR == 2 ? l_part[row] + l_part[row + 128] : l_run;
            {
                const float inv = (my_seg != 0.f && l_tot > 0.f) ? 1.0f / l_tot : 0.f;
                const uint32_t dst = stage_sm + row * (DH * 2) + half * (OH * 2);
#pragma unroll
                for (int c = 0; c < OH / 8; ++c) {
                    if (half * OH + c * 8 < DH) {
                        uint32_t pk[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) pk[p] = Op16<OT>::pack(o[c * 8 + 2 * p] * inv, o[c * 8 + 2 * p + 1] * inv);
                        sts128(dst + c * 16, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    }
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(1 + x, SOFT_THREADS);
            if (tid_wg == 0) tma_store_4d(&map_o, stage_sm, 0, head, qt * 128, sample);
            store_pending = true;
            ++n_item;
        }
        if (store_pending && tid_wg == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace fitv2
