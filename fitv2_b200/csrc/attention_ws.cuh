// Warp-specialised, software-pipelined masked flash attention (production kernel).
//
// Reference: fit/model/modules.py:176-204 (segment-id mask, softmax, P V, output * (mask != 0)).  Same operand layouts as
// attention_general.cuh, but a bound-based single-pass softmax (no running maximum: q and k are affine-free-LayerNorm'ed)
// and a different schedule.  attention_general.cuh runs  S-MMA -> softmax -> PV-MMA  strictly in sequence inside a CTA
// (ncu: 28 % issue slots, stalls on barriers / TMEM loads, MUFU 32 %); here one persistent CTA per SM keeps TWO query tiles
// ("streams" a and b, the two 128-row tiles of one (sample, head) pair at 256 tokens) in flight:
//
//   warp 0        TMA producer: Q_a, Q_b once per work item, K / V^T tiles through 2-stage rings; it runs ahead
//                 of the consumers, so the loads of the next item hide behind the current one.
//   warps 1, 2    tcgen05 issuers, one per stream: S_x = Q_x K_t^T into a per-stream TMEM buffer, O_x += P_x V_t.
//                 The next S_x (of this or of the NEXT work item) is issued as soon as the softmax warps have READ
//                 the previous one out of TMEM (s_free), i.e. it overlaps the exponentials; P_x V_t is issued when
//                 P_x(t) has been written (p_full).  One issuer per stream, so a stream never waits for the other
//                 (a single issuer walking both streams in a fixed order cost 58 us, a polling one 63 us).
//   warps 3-18    softmax: one 8-warp group per stream (two threads per query row, 64 key columns each).  Measured
//                 alternatives (profiles/README.md): all 16 warps on one tile alternating between the streams (52.9 us at 256
//                 tokens against 45.8: every barrier round trip becomes serial); exp2 on the FMA pipe for a fraction of the
//                 elements, dropping the scale FMA / the row-sum adds, starting stream b late, a CTA-pair variant with P and a
//                 double-buffered S in TMEM (round 2, all slower or equal: the XU pipe sits at 59-67 % and the rest of the
//                 time is the chain of MIO round trips -- barrier test, TMEM load, TMEM / shared store, arrive -- per tile).
//                 The normalised output tile leaves through shared memory and ONE TMA store per tile
//                 (row-per-thread global stores cost ~2000 clk per tile).
//
// TMEM: S_a [0,128) S_b [128,256) O_a [256,256+DHP) O_b [384,384+DHP).  Shared memory (head_dim 72): Q 2x20 KB,
// K ring 2x20 KB, V^T ring 2x20 KB, P 2x32 KB = 184 KB (+ segment ids of the sample for the masked path).
#pragma once
#include "common.cuh"
#include "tc2sm.cuh"

namespace fitv2 {

template <int DH> struct AttnWsCfg {
    static constexpr int kDHP = (DH + 15) / 16 * 16;          // 72 -> 80, 96 -> 96
    static constexpr int kTail = kDHP - 64;                    // elements in the tail panel (16 or 32)
    static constexpr int kTailBytes = kTail * 2;               // 32 or 64
    static constexpr int kQMain = 128 * 128;
    static constexpr int kQTail = 128 * kTailBytes;
    static constexpr int kQKTile = kQMain + ((kQTail + 1023) / 1024) * 1024;   // one Q or K tile (main + tail panel)
    static constexpr int kVPanel = ((kDHP * 128 + 1023) / 1024) * 1024;        // one 64-key panel of V^T
    static constexpr int kVTile = 2 * kVPanel;
    static constexpr int kPPanel = 128 * 128;
    static constexpr int kPTile = 2 * kPPanel;
    static constexpr int kKStages = 2, kVStages = 2;
    static constexpr int kOffQ = 0;
    static constexpr int kOffK = kOffQ + 2 * kQKTile;
    static constexpr int kOffV = kOffK + kKStages * kQKTile;
    static constexpr int kOffP = kOffV + kVStages * kVTile;
    static constexpr int kOffSum = kOffP + 2 * kPTile;         // 2 streams x 4 column quarters x 128 partial row sums
    static constexpr int kOffBar = kOffSum + 2 * 512 * 4;
    static constexpr int kNumBars = 8 + 2 * kKStages + 2 * kVStages + 10;
    static constexpr int kOffSeg = kOffBar + ((kNumBars * 8 + 16 + 127) / 128) * 128;   // 2 x seg_pad floats (dynamic)
    static constexpr uint32_t kQKBytes = kQMain + kQTail;      // TMA transaction bytes of one Q or K tile
    static constexpr uint32_t kVBytes = 2 * kDHP * 128;
    static constexpr int kThreads = 32 * 19;
    static constexpr int smem_bytes(int tokens) { return kOffSeg + 2 * ((tokens + 127) / 128 * 128) * 4 + 1024; }
    static_assert(kTail == 16 || kTail == 32, "head_dim must be 72..80 or 88..96 (64 + 16/32 tail)");
};

#ifdef FITV2_ATTN_TRACE
// timing experiments only: per-role event timestamps of CTA 0 (role, event index) -> (tag, clock64)
__device__ unsigned long long g_attn_trace[20][512][2];
__device__ unsigned int g_attn_trace_n[20];
#define ATTN_TRACE(role, tag) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { unsigned int i_ = tr_i++; \
    if (i_ < 512) { g_attn_trace[role][i_][0] = (unsigned long long)(tag); g_attn_trace[role][i_][1] = clock64(); g_attn_trace_n[role] = i_ + 1; } } } while (0)
#else
#define ATTN_TRACE(role, tag) do { } while (0)
#endif

// 4-D TMA tile store shared -> global (bulk async group); rows / columns outside the tensor are clipped.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

template <typename OT, int DH>
__global__ void __launch_bounds__(608, 1)
attention_ws_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qt,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_kt,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                    const float* __restrict__ seg, const int* __restrict__ seg_uniform,
                    int heads, int tokens, int num_items, float scale_log2e, float bound_log2e)
{
    using C = AttnWsCfg<DH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;                    // [2]  Q_x landed
    uint64_t* q_empty = bars + 2;               // [2]  every S MMA of the item that reads Q_x has retired
    uint64_t* k_full = bars + 4;                // [kKStages]
    uint64_t* k_empty = k_full + C::kKStages;   //      both streams have released the stage (2 arrivals)
    uint64_t* v_full = k_empty + C::kKStages;   // [kVStages]
    uint64_t* v_empty = v_full + C::kVStages;
    uint64_t* s_full = v_empty + C::kVStages;   // [2]  S_x(t) accumulated in TMEM
    uint64_t* s_free = s_full + 2;              // [2]  softmax warps have read S_x(t) (16 warp arrivals)
    uint64_t* p_full = s_free + 2;              // [2]  P_x(t) written to shared memory (16 warp arrivals)
    uint64_t* pv_done = p_full + 2;             // [2]  O_x += P_x(t) V_t retired: P_x buffer free; after the last tile O_x is final
    uint64_t* o_free = pv_done + 2;             // [2]  softmax warps have read O_x (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tiles = (tokens + 127) / 128, kv_tiles = q_tiles;
    const int q_pairs = (q_tiles + 1) / 2;
#ifdef FITV2_ATTN_TRACE
    unsigned int tr_i = 0;
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_qt); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_kt);
        tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 8);
            mbar_init(&p_full[i], 8); mbar_init(&pv_done[i], 1); mbar_init(&o_free[i], 8);
        }
        for (int i = 0; i < C::kKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }
        for (int i = 0; i < C::kVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                 // Q / K / V^T come from the QKV GEMM in front of this kernel
    pdl_launch_dependents();

    if (warp == 0) {
        // ------------------------------------- TMA producer -------------------------------------
        // Issue order = the order in which the buffers become free, not the order of use: the first K tile of the
        // NEXT item goes out before the last V tile of this one (its stage is released by the first S tiles, early),
        // and K(t+1) before V(t).  With the straightforward Q, K0, V0, K1, V1 order every item start waited a full
        // TMA latency (~2400 clk of 11 000 per item) for K0, whatever the schedule of the consumers.
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
            ATTN_TRACE(0, 110 + t);
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
            ATTN_TRACE(0, 120 + t);
            if (++vs == C::kVStages) { vs = 0; vph ^= 1; }
        };
        if ((int)blockIdx.x < num_items) load_k(blockIdx.x / q_pairs, 0);
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int bh = item / q_pairs, qp = item - bh * q_pairs;
            const int nstreams = (2 * qp + 1 < q_tiles) ? 2 : 1;
            for (int x = 0; x < nstreams; ++x) {
                if (n_q[x] > 0) mbar_wait(&q_empty[x], (n_q[x] - 1) & 1);
                if (elect_one()) {
                    uint8_t* dst = smem + C::kOffQ + x * C::kQKTile;
                    mbar_arrive_expect_tx(&q_full[x], C::kQKBytes);
                    tma_load_3d(&map_q, &q_full[x], dst, 0, (2 * qp + x) * 128, bh);
                    tma_load_3d(&map_qt, &q_full[x], dst + C::kQMain, 64, (2 * qp + x) * 128, bh);
                }
                __syncwarp();
                ATTN_TRACE(0, 100 + x);
                ++n_q[x];
            }
            const int next_item = item + gridDim.x;
            for (int t = 0; t < kv_tiles; ++t) {
                if (t + 1 < kv_tiles) load_k(bh, t + 1);
                else if (next_item < num_items) load_k(next_item / q_pairs, 0);
                load_v(bh, t);
            }
        }
    } else if (warp <= 2) {
        // ------------------------------------- tcgen05 issuer of stream x -------------------------------------
        // Two in-order sequences, the S tiles and the P V products of the stream; S runs one tile ahead (across
        // work items), so the order is  S(0) | S(1) PV(0) | S(2) PV(1) | ... : each wait is for an event that
        // does not depend on anything issued later.  A stream that idles in an item (odd tile count) only hands
        // its share of the K / V stages back.
        const int x = warp - 1;
        constexpr uint32_t idesc_s = umma_idesc(Op16<OT>::kUmmaFormat, 128, 128);
        constexpr uint32_t idesc_o = umma_idesc(Op16<OT>::kUmmaFormat, 128, C::kDHP);
        const uint32_t sm_q = smem_u32(smem + C::kOffQ) + x * C::kQKTile, sm_k = smem_u32(smem + C::kOffK);
        const uint32_t sm_v = smem_u32(smem + C::kOffV), sm_p = smem_u32(smem + C::kOffP) + x * C::kPTile;
        const uint32_t d_s = tmem_base + x * 128, d_o = tmem_base + 256 + x * 128;
        auto active = [&](int item) { return x == 0 || 2 * (item % q_pairs) + 1 < q_tiles; };
        struct Seq { int item, t; uint32_t n, n_item, stage, phase; };
        Seq sq = {(int)blockIdx.x, 0, 0u, 0u, 0u, 0u}, pv = sq;
        auto s_step = [&]() {                                           // issue the next S tile of this stream (if any)
            if (sq.item < num_items) {
                mbar_wait(&k_full[sq.stage], sq.phase);
                const bool act = active(sq.item);
                if (act) {
                    if (sq.n > 0) mbar_wait(&s_free[x], (sq.n - 1) & 1);
                    if (sq.t == 0) mbar_wait(&q_full[x], sq.n_item & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t dq = umma_desc_kmajor(sm_q, 128);
                        const uint64_t dk = umma_desc_kmajor(sm_k + sq.stage * C::kQKTile, 128);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_ss(d_s, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
                        const uint64_t dqt = umma_desc_kmajor(sm_q + C::kQMain, C::kTailBytes);
                        const uint64_t dkt = umma_desc_kmajor(sm_k + sq.stage * C::kQKTile + C::kQMain, C::kTailBytes);
#pragma unroll
                        for (int kk = 0; kk < C::kTail / 16; ++kk) umma_ss(d_s, dqt + 2 * kk, dkt + 2 * kk, idesc_s, 1);
                        umma_commit(&s_full[x]);
                        umma_commit(&k_empty[sq.stage]);
                        if (sq.t + 1 == kv_tiles) umma_commit(&q_empty[x]);
                    }
                    __syncwarp();
                    ATTN_TRACE(warp, 200 + sq.t);
                    ++sq.n;
                } else {
                    if (elect_one()) mbar_arrive(&k_empty[sq.stage]);
                    __syncwarp();
                }
                if (++sq.stage == C::kKStages) { sq.stage = 0; sq.phase ^= 1; }
                if (++sq.t == kv_tiles) { sq.t = 0; sq.item += gridDim.x; if (act) ++sq.n_item; }
                return;                                                 // one tile per call, also for an idle stream: K and V stages must be
            }                                                           // handed back in the producer's order or the rings deadlock
        };
        auto pv_step = [&]() {
            if (pv.item < num_items) {
                mbar_wait(&v_full[pv.stage], pv.phase);
                const bool act = active(pv.item);
                if (act) {
                    mbar_wait(&p_full[x], pv.n & 1);
                    if (pv.t == 0 && pv.n_item > 0) mbar_wait(&o_free[x], (pv.n_item - 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint64_t dp = umma_desc_kmajor(sm_p + (kk >> 2) * C::kPPanel, 128) + 2 * (kk & 3);
                            const uint64_t dv = umma_desc_kmajor(sm_v + pv.stage * C::kVTile + (kk >> 2) * C::kVPanel, 128) + 2 * (kk & 3);
                            umma_ss(d_o, dp, dv, idesc_o, (pv.t | kk) != 0);
                        }
                        umma_commit(&pv_done[x]);
                        umma_commit(&v_empty[pv.stage]);
                    }
                    __syncwarp();
                    ATTN_TRACE(warp, 300 + pv.t);
                    ++pv.n;
                } else {
                    if (elect_one()) mbar_arrive(&v_empty[pv.stage]);
                    __syncwarp();
                }
                if (++pv.stage == C::kVStages) { pv.stage = 0; pv.phase ^= 1; }
                if (++pv.t == kv_tiles) { pv.t = 0; pv.item += gridDim.x; if (act) ++pv.n_item; }
                return;                                                 // one tile per call, also for an idle stream: K and V stages must be
            }                                                           // handed back in the producer's order or the rings deadlock
        };
        s_step();
        while (pv.item < num_items) { s_step(); pv_step(); }
    } else {
        // ------------------------------------- softmax warp groups -------------------------------------
        const int x = (warp - 3) >> 3;                                  // stream: 0 = first query tile of the pair, 1 = second
        const int wl = (warp - 3) & 7;
        const int half = wl >> 2;                                       // which 64 key columns / output half this thread owns
        const int quarter = warp & 3;                                   // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;                            // query row inside the tile == TMEM lane
        const int tid_wg = wl * 32 + lane;
        const uint32_t t_s = tmem_base + (uint32_t(quarter * 32) << 16) + x * 128 + half * 64;
        const uint32_t t_o = tmem_base + (uint32_t(quarter * 32) << 16) + 256 + x * 128;
        const uint32_t smem_px = smem_u32(smem + C::kOffP) + x * C::kPTile;   // P_x; reused as the output staging tile
        const uint32_t smem_p = smem_px + half * C::kPPanel;
        float* l_part = reinterpret_cast<float*>(smem + C::kOffSum) + x * 256;
        const int seg_pad = (tokens + 127) / 128 * 128;
        float* seg_s = reinterpret_cast<float*>(smem + C::kOffSeg) + x * seg_pad;
        uint32_t n_s = 0;
        bool store_pending = false;                                     // a TMA store may still be reading the staging tile
        // per-item scalars are fetched one item ahead (two dependent global loads would otherwise sit on the item start)
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
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int klen = uni_nx;                                    // key length of the sample (seg_uniform_kernel), 0 = compare ids
            const bool uniform = klen != 0;
            const float my_seg = seg_nx;
            fetch_meta(item + gridDim.x);
            const int bh = item / q_pairs, qp = item - bh * q_pairs;
            const int qt = 2 * qp + x;
            if (qt >= q_tiles) continue;                                // odd number of query tiles: stream b idles
            const int sample = bh / heads, head = bh - sample * heads;
            if (!uniform) {                                             // masked path: key segment ids of the sample in smem
                // (the readers of the previous item's ids have all passed that item's epilogue barriers)
                const float* segb = seg + (size_t)sample * tokens;
                for (int i = tid_wg; i < seg_pad; i += 256) seg_s[i] = i < tokens ? __ldg(segb + i) : 0.f;
                named_bar_sync(1 + x, 256);
            }
            float l_run = 0.f;
            for (int t = 0; t < kv_tiles; ++t, ++n_s) {
                const int kv0 = t * 128;
                const int kv_valid = min(128, (uniform ? klen : tokens) - kv0);   // may be <= 0 behind the last valid key
                const int mode = (uniform && kv_valid == 128) ? 0 : (uniform ? 1 : 2);   // dense / key-tail bound / segment compare
                ATTN_TRACE(warp, 400 + t);
                mbar_wait(&s_full[x], n_s & 1);
                tc_fence_after();
                ATTN_TRACE(warp, 410 + t);
                uint32_t packed[32];
                float lsum = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32(t_s + c * 32, v);
                    tmem_ld_wait();
                    if (c == 1) {                                       // S_x is in registers: the next Q K^T may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s_free[x]);
                        ATTN_TRACE(warp, 420 + t);
                    }
                    auto soft32 = [&](auto mode_c) {                    // mode is a compile-time constant inside
                        constexpr int kMode = decltype(mode_c)::value;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -bound_log2e));
                            float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -bound_log2e));
                            if constexpr (kMode != 0) {
                                const int col = half * 64 + c * 32 + 2 * j;
                                bool ok0 = col < kv_valid, ok1 = col + 1 < kv_valid;
                                if constexpr (kMode == 2) {
                                    ok0 = ok0 && seg_s[kv0 + col] == my_seg;
                                    ok1 = ok1 && seg_s[kv0 + col + 1] == my_seg;
                                }
                                p0 = ok0 ? p0 : 0.f;
                                p1 = ok1 ? p1 : 0.f;
                            }
                            packed[c * 16 + j] = Op16<OT>::pack(p0, p1);
                            lsum += p0 + p1;
                        }
                    };
                    if (mode == 0) soft32(std::integral_constant<int, 0>{});
                    else if (mode == 1) soft32(std::integral_constant<int, 1>{});
                    else soft32(std::integral_constant<int, 2>{});
                }
                l_run += lsum;
                ATTN_TRACE(warp, 430 + t);
                if (n_s > 0) mbar_wait(&pv_done[x], (n_s - 1) & 1);     // the previous P V of this stream has read the P buffer
                if (t == 0 && store_pending) {                          // ... and so has the TMA store of the previous output tile
                    if (tid_wg == 0) tma_store_wait_read();
                    named_bar_sync(1 + x, 256);
                    store_pending = false;
                }
                ATTN_TRACE(warp, 440 + t);
#pragma unroll
                for (int g = 0; g < 8; ++g)                             // 8 chunks of 8 keys: this thread's 64-key panel row
                    sts128(smem_p + swz_offset<128>(row, g),
                           make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[x]);
                ATTN_TRACE(warp, 450 + t);
            }

            // ---- O / l, zero padded queries (mask != 0): (128, DH) tile -> staging -> one TMA store into the
            //      (M, heads*DH) rows that feed the proj GEMM ----
            l_part[half * 128 + row] = l_run;
            mbar_wait(&pv_done[x], (n_s - 1) & 1);                      // last P V retired: O_x is final, P_x is free
            tc_fence_after();
            ATTN_TRACE(warp, 500);
            constexpr int OH = C::kDHP / 2;                            // output columns per thread: 40 or 48
            float o[OH];
            tmem_ld32(t_o + half * OH, reinterpret_cast<uint32_t*>(o));
            if constexpr (OH == 40) tmem_ld8(t_o + half * OH + 32, reinterpret_cast<uint32_t*>(o) + 32);
            else tmem_ld16(t_o + half * OH + 32, reinterpret_cast<uint32_t*>(o) + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[x]);
            ATTN_TRACE(warp, 510);
            named_bar_sync(1 + x, 256);                                 // both halves' row sums are in l_part
            const float l_tot = l_part[row] + l_part[row + 128];
            ATTN_TRACE(warp, 520);
            {
                const float inv = (my_seg != 0.f && l_tot > 0.f) ? 1.0f / l_tot : 0.f;
                const uint32_t dst = smem_px + row * (DH * 2) + half * (OH * 2);
#pragma unroll
                for (int c = 0; c < OH / 8; ++c) {
                    if (half * OH + c * 8 < DH) {                      // skip the zero-pad columns 72..79
                        uint32_t pk[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) pk[p] = Op16<OT>::pack(o[c * 8 + 2 * p] * inv, o[c * 8 + 2 * p + 1] * inv);
                        sts128(dst + c * 16, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                    }
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(1 + x, 256);
            if (tid_wg == 0) tma_store_4d(&map_o, smem_px, 0, head, qt * 128, sample);   // rows >= tokens are clipped
            store_pending = true;
            ATTN_TRACE(warp, 530);
        }
        if (store_pending && tid_wg == 0) tma_store_wait_all();         // global writes complete before the CTA retires
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace fitv2
