// Masked flash attention on CTA PAIRS (tcgen05 cta_group::2): the production kernel for the bound-based softmax.
//
// Same math, reference lines and operand layouts in HBM as attention_ws.cuh (fit/model/modules.py:176-204; constant logit
// bound instead of a running maximum).  What changes is the decomposition.  attention_ws / attention_tm keep two query tiles
// ("streams") of one (sample, head) in ONE CTA; measured on B200 (profiles/r2_attn_*.log, timelines r2_attn_trace_1024*.txt)
// they are bound neither by MUFU throughput (exp2 on the FMA pipe: slower) nor by issue slots (no FMA / no row sum: same time)
// but by the per-key-tile synchronisation chain (S wait -> TMEM load -> exp -> P store -> proxy fence -> arrive -> P V) that
// all softmax warps walk in lockstep, and by shared-memory bandwidth (293 KB of operand reads / P round trips / TMA writes
// per 2 x 128 x 128 tile pair = 2290 clk at 128 B/clk against 2048 clk of exponentials).  Here:
//
//   * a CLUSTER OF TWO CTAs owns the two 128-row query tiles of one (sample, head): one tile per SM, all 16 softmax warps of
//     an SM on the same tile (4 threads per query row, 32 key columns each);
//   * both contractions are M = 256 UMMAs issued by the leader CTA (cta_group::2): S = [Q_0; Q_1] K_t^T with each CTA staging
//     only HALF of the key tile (64 keys), O += P V_t with each CTA staging half of the V^T rows; TMA bytes and UMMA operand
//     reads per SM drop to 60 KB per key tile (470 clk), so the exponentials (1024 clk per tile) are the only bound left;
//   * TMEM of one SM now has room for everything a decoupled pipeline needs:  S double-buffered (2 x 128 columns), O (80 / 96),
//     P double-buffered in TENSOR MEMORY (2 x 64 columns of packed 16-bit pairs, A operand of P V: no shared-memory round trip,
//     no proxy fence).  S(t+2) only waits for the softmax warps to have READ S(t); P(t) only for P V(t-2): the softmax warps
//     never wait for the tensor pipe in steady state and drift out of phase instead of marching in lockstep;
//   * the first key tile of the NEXT work item is processed before the epilogue of the current one (its S is ready: S runs two
//     tiles ahead across items), which hides the drain of the last P V.
//
//   warp 0      TMA producer (every CTA: its Q tile, its halves of K_t / V_t^T; 4-stage rings; bytes credited to the leader)
//   warp 1      tcgen05 issuer (leader CTA only) + TMEM allocation
//   warps 2-17  softmax + epilogue (normalise, stage, ONE TMA store per 128 x head_dim tile)
//
// TMEM columns per CTA: S0 [0,128) S1 [128,256) O [256,256+DHP) P0 [256+DHP, +64) P1 [.., +64)  (464 at head_dim 72, 480 at 96).
#pragma once
#include "attention_tm.cuh"

namespace fitv2 {

template <int DH> struct AttnP2Cfg {
    using W = AttnWsCfg<DH>;
    static constexpr int kDHP = W::kDHP, kTail = W::kTail, kTailBytes = W::kTailBytes;
    static constexpr int kQMain = W::kQMain, kQKTile = W::kQKTile;           // one 128-row Q tile (main + tail panel)
    static constexpr int kKMain = 64 * 128;                                  // 64 key rows of the 64-element main panel
    static constexpr int kKTile = kKMain + ((64 * kTailBytes + 1023) / 1024) * 1024;
    static constexpr int kVRows = kDHP / 2;                                  // V^T rows (head-dim index) staged per CTA
    static constexpr int kVPanel = ((kVRows * 128 + 1023) / 1024) * 1024;    // one 64-key panel of those rows
    static constexpr int kVTile = 2 * kVPanel;
    static constexpr int kStages = 4;
    static constexpr int kOffQ = 0;                                          // [2] double-buffered across work items
    static constexpr int kOffK = kOffQ + 2 * kQKTile;
    static constexpr int kOffV = kOffK + kStages * kKTile;
    static constexpr int kOffStage = kOffV + kStages * kVTile;               // output staging tile (dense rows of DH elements)
    static constexpr int kStageTile = ((128 * DH * 2 + 1023) / 1024) * 1024;
    static constexpr int kOffSum = kOffStage + kStageTile;                   // 4 column quarters x 128 partial row sums
    static constexpr int kOffBar = kOffSum + 4 * 128 * 4;
    static constexpr int kNumBars = 4 + 4 * kStages + 12;
    static constexpr int kOffSeg = kOffBar + ((kNumBars * 8 + 16 + 127) / 128) * 128;
    static constexpr uint32_t kQBytes = W::kQKBytes;                         // TMA bytes of one Q tile
    static constexpr uint32_t kKBytes = 64 * 128 + 64 * kTailBytes;          // ... of one half key tile
    static constexpr uint32_t kVBytes = 2 * kVRows * 128;                    // ... of one half V^T tile
    static constexpr int kColS = 0, kColO = 256, kColP = 256 + kDHP;
    static constexpr int kThreads = 32 * 18;
    static constexpr int smem_bytes(int tokens) { return kOffSeg + ((tokens + 127) / 128 * 128) * 4 + 1024; }
    static_assert(kColP + 128 <= 512, "TMEM budget");
    static_assert(kVRows % 8 == 0, "V^T half must be whole 8-row swizzle atoms");
};

// 3-D TMA tile load executed by either CTA of the pair; bytes credited to the LEADER's mbarrier (see tma_load_2d_2sm).
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// D[tmem of both CTAs] (+)= A[tmem, 128 lanes per CTA] * B[smem, N/2 rows per CTA], M = 256.  Leader CTA, one thread.
__device__ __forceinline__ void umma_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on the LEADER CTA's copy of `bar` from either CTA of the pair.  What the barrier orders are TMEM accesses
// (tcgen05.wait::ld / ::st + tcgen05.fence::before_thread_sync in front of the arrive, tcgen05.fence::after_thread_sync behind the
// wait), not generic memory, so the default CTA-scope semantics suffice -- a cluster-scope release costs > 1000 clk per arrive
// (measured: profiles/r2_attn_trace_p2_release_cluster.txt) and sat on the softmax warps' critical path twice per key tile.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar, bool leader) {
    if (leader) mbar_arrive(bar);
    else mbar_arrive_remote(bar, 0);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
// tcgen05.wait::ld that also "touches" the 16 destination registers of an earlier tcgen05.ld: the loads are asynchronous and the
// wait has no data dependence on them, so without this the compiler would be free to move arithmetic on the registers above
// the wait when the load and its use are far apart (software-pipelined prefetch).
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(addr), "r"(a), "r"(b) : "memory");
}

template <typename OT, int DH>
__global__ void __launch_bounds__(576, 1)
attention_p2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qt,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_kt,     // 64-row boxes
                    const __grid_constant__ CUtensorMap map_v,                                                    // (64 keys, DHP/2 rows) boxes
                    const __grid_constant__ CUtensorMap map_o,
                    const float* __restrict__ seg, const int* __restrict__ seg_uniform,
                    int heads, int tokens, int num_items, float scale_log2e, float bound_log2e)
{
    using C = AttnP2Cfg<DH>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;                        // [2] leader: both CTAs' Q tiles landed
    uint64_t* q_empty = bars + 2;                   // [2] every CTA: the S MMAs of the item have retired (multicast commit)
    uint64_t* k_full = bars + 4;                    // [kStages] leader
    uint64_t* k_empty = k_full + C::kStages;        // [kStages] every CTA
    uint64_t* v_full = k_empty + C::kStages;        // [kStages] leader
    uint64_t* v_empty = v_full + C::kStages;        // [kStages] every CTA
    uint64_t* s_full = v_empty + C::kStages;        // [2] every CTA: S(t) accumulated in TMEM buffer t % 2
    uint64_t* s_free = s_full + 2;                  // [2] leader: all 32 softmax warps of the pair hold S(t) in registers
    uint64_t* p_full = s_free + 2;                  // [2] leader: all 32 softmax warps have written P(t)
    uint64_t* pv_done = p_full + 2;                 // [2] every CTA: P V(t) retired, P buffer t % 2 is free
    uint64_t* o_full = pv_done + 2;                 // [1] every CTA: last P V of the item retired
    uint64_t* o_free = o_full + 1;                  // [1] leader: all 32 softmax warps have read O
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int q_tiles = (tokens + 127) / 128, kv_tiles = q_tiles;
    const int q_pairs = (q_tiles + 1) / 2;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
#ifdef FITV2_ATTN_TRACE
    unsigned int tr_i = 0;
#endif

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_qt); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_kt);
        tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 32);
            mbar_init(&p_full[i], 32); mbar_init(&pv_done[i], 1);
        }
        mbar_init(o_full, 1); mbar_init(o_free, 32);
        for (int i = 0; i < C::kStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
    tc_fence_before();
    cluster_sync();                                 // the peer's barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                     // Q / K / V^T come from the QKV GEMM in front of this kernel
    pdl_launch_dependents();

    if (warp == 0) {
        // ------------------------------------- TMA producer (every CTA) -------------------------------------
        uint32_t ks = 0, kph = 0, vs = 0, vph = 0, n_item = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters, ++n_item) {
            const int bh = item / q_pairs, qp = item - bh * q_pairs;
            const uint32_t qb = n_item & 1;
            if (n_item >= 2) mbar_wait(&q_empty[qb], ((n_item >> 1) - 1) & 1);
            if (elect_one()) {
                uint8_t* dst = smem + C::kOffQ + qb * C::kQKTile;
                if (leader) mbar_arrive_expect_tx(&q_full[qb], 2 * C::kQBytes);
                tma_load_3d_2sm(&map_q, &q_full[qb], dst, 0, (2 * qp + (int)rank) * 128, bh);       // rows past the sequence: zero fill
                tma_load_3d_2sm(&map_qt, &q_full[qb], dst + C::kQMain, 64, (2 * qp + (int)rank) * 128, bh);
            }
            __syncwarp();
            for (int t = 0; t < kv_tiles; ++t) {
                mbar_wait(&k_empty[ks], kph ^ 1);
                if (elect_one()) {
                    uint8_t* dst = smem + C::kOffK + ks * C::kKTile;
                    if (leader) mbar_arrive_expect_tx(&k_full[ks], 2 * C::kKBytes);
                    tma_load_3d_2sm(&map_k, &k_full[ks], dst, 0, t * 128 + (int)rank * 64, bh);
                    tma_load_3d_2sm(&map_kt, &k_full[ks], dst + C::kKMain, 64, t * 128 + (int)rank * 64, bh);
                }
                __syncwarp();
                if (++ks == C::kStages) { ks = 0; kph ^= 1; }
                mbar_wait(&v_empty[vs], vph ^ 1);
                if (elect_one()) {
                    uint8_t* dst = smem + C::kOffV + vs * C::kVTile;
                    if (leader) mbar_arrive_expect_tx(&v_full[vs], 2 * C::kVBytes);
                    tma_load_3d_2sm(&map_v, &v_full[vs], dst, t * 128, (int)rank * C::kVRows, bh);
                    tma_load_3d_2sm(&map_v, &v_full[vs], dst + C::kVPanel, t * 128 + 64, (int)rank * C::kVRows, bh);
                }
                __syncwarp();
                if (++vs == C::kStages) { vs = 0; vph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ------------------------------------- tcgen05 issuer (leader CTA) -------------------------------------
            // Two in-order sequences over the global tile counter n: S(n) and P V(n).  S runs two tiles ahead (also across work
            // items): the order is  S(0) S(1) | S(2) PV(0) | S(3) PV(1) | ...   S(n+2) waits for the softmax warps to have READ
            // S(n) (early in their tile), P V(n) for P(n) (late), so every wait is for an event that does not depend on anything
            // issued after it.
            constexpr uint32_t idesc_s = umma_idesc(Op16<OT>::kUmmaFormat, 256, 128);
            constexpr uint32_t idesc_o = umma_idesc(Op16<OT>::kUmmaFormat, 256, C::kDHP);
            const uint32_t sm_q = smem_u32(smem + C::kOffQ), sm_k = smem_u32(smem + C::kOffK), sm_v = smem_u32(smem + C::kOffV);
            struct Seq { int item, t; uint32_t n, n_item, stage, phase; };
            Seq sq = {cluster_id, 0, 0u, 0u, 0u, 0u}, pv = sq;
            auto s_step = [&]() {
                if (sq.item >= num_items) return;
                const uint32_t b = sq.n & 1, qb = sq.n_item & 1;
                mbar_wait(&k_full[sq.stage], sq.phase);
                if (sq.t == 0) mbar_wait(&q_full[qb], (sq.n_item >> 1) & 1);
                if (sq.n >= 2) mbar_wait(&s_free[b], ((sq.n >> 1) - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + C::kColS + b * 128;
                    const uint64_t dq = umma_desc_kmajor(sm_q + qb * C::kQKTile, 128);
                    const uint64_t dk = umma_desc_kmajor(sm_k + sq.stage * C::kKTile, 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_ss_2sm(d, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
                    const uint64_t dqt = umma_desc_kmajor(sm_q + qb * C::kQKTile + C::kQMain, C::kTailBytes);
                    const uint64_t dkt = umma_desc_kmajor(sm_k + sq.stage * C::kKTile + C::kKMain, C::kTailBytes);
#pragma unroll
                    for (int kk = 0; kk < C::kTail / 16; ++kk) umma_ss_2sm(d, dqt + 2 * kk, dkt + 2 * kk, idesc_s, 1);
                    umma_commit_2sm(&s_full[b]);
                    umma_commit_2sm(&k_empty[sq.stage]);
                    if (sq.t + 1 == kv_tiles) umma_commit_2sm(&q_empty[qb]);
                }
                __syncwarp();
                ATTN_TRACE(1, 200 + sq.t);
                ++sq.n;
                if (++sq.stage == C::kStages) { sq.stage = 0; sq.phase ^= 1; }
                if (++sq.t == kv_tiles) { sq.t = 0; sq.item += num_clusters; ++sq.n_item; }
            };
            auto pv_step = [&]() {
                const uint32_t b = pv.n & 1;
                mbar_wait(&v_full[pv.stage], pv.phase);
                mbar_wait(&p_full[b], (pv.n >> 1) & 1);
                if (pv.t == 0 && pv.n_item > 0) mbar_wait(o_free, (pv.n_item - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a = tmem_base + C::kColP + b * 64, d = tmem_base + C::kColO;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {                        // 16 keys per step: 8 packed columns of P, 32 bytes of a V^T row
                        const uint64_t dv = umma_desc_kmajor(sm_v + pv.stage * C::kVTile + (kk >> 2) * C::kVPanel, 128) + 2 * (kk & 3);
                        umma_ts_2sm(d, a + kk * 8, dv, idesc_o, (pv.t | kk) != 0);
                    }
                    umma_commit_2sm(&pv_done[b]);
                    umma_commit_2sm(&v_empty[pv.stage]);
                    if (pv.t + 1 == kv_tiles) umma_commit_2sm(o_full);
                }
                __syncwarp();
                ATTN_TRACE(1, 300 + pv.t);
                ++pv.n;
                if (++pv.stage == C::kStages) { pv.stage = 0; pv.phase ^= 1; }
                if (++pv.t == kv_tiles) { pv.t = 0; pv.item += num_clusters; ++pv.n_item; }
            };
            s_step(); s_step();
            while (pv.item < num_items) { s_step(); pv_step(); }
        }
    } else {
        // ------------------------------------- softmax + epilogue warps (every CTA: its own 128 query rows) -------------------------------------
        const int sw = warp - 2;                                            // 0..15
        const int cq = sw >> 2;                                             // column quarter: 32 keys of every tile / a quarter of the output columns
        const int quarter = warp & 3;                                       // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;                                // query row inside the tile == TMEM lane
        const int tid_sm = sw * 32 + lane;                                  // 0..511
        const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16);
        const uint32_t stage_sm = smem_u32(smem + C::kOffStage);
        float* l_part = reinterpret_cast<float*>(smem + C::kOffSum);
        const int seg_pad = (tokens + 127) / 128 * 128;
        float* seg_s = reinterpret_cast<float*>(smem + C::kOffSeg);
        uint32_t n_s = 0, n_item = 0;
        bool store_pending = false;
        int uni_nx = 1; float seg_nx = 0.f;
        auto fetch_meta = [&](int item) {
            if (item < num_items) {
                const int bh = item / q_pairs, qp = item - bh * q_pairs;
                const int sample = bh / heads, qi = (2 * qp + (int)rank) * 128 + row;
                uni_nx = __ldg(seg_uniform + sample);
                seg_nx = qi < tokens ? __ldg(seg + (size_t)sample * tokens + qi) : 0.f;
            }
        };
        fetch_meta(cluster_id);
        // One 128-key tile of the softmax, software-pipelined inside the warp.  Every softmax warp walks the same chain per tile
        // (barrier test -> TMEM load -> exponentials -> TMEM store -> arrive) and the four warps of a scheduler stay in phase, so
        // each latency in the chain is a bubble in the MUFU (measured: ~830 clk of serial latency against 1024 clk of
        // exponentials per tile, profiles/r2_attn_trace_p2.txt).  Here the latencies are taken off the chain instead:
        //   * the thread's 32 columns of S(n) are two 16-column chunks A / B; chunk A of tile n+1 is requested from TMEM in
        //     the middle of tile n, chunk B of tile n at its start, each a half tile of exponentials before it is needed;
        //   * the barrier tests of the next tile (s_full) and of the P buffer (pv_done) are ISSUED before the exponentials and
        //     only looked at after them;
        //   * the TMEM store of P(n) is issued at the end of tile n and waited for / signalled (p_full) in the middle of tile n+1.
        // regA holds chunk A of the tile about to be processed (a_loaded), `pend` says a P store still has to be signalled.
        uint32_t regA[16];
        bool a_loaded = false, pend = false;
        auto flush_p = [&](uint32_t buf) {                                  // P(buf) store complete -> p_full
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&p_full[buf], leader);
            pend = false;
        };
        auto do_tile = [&](int t, bool uniform, float my_seg, bool prefetch_next) -> float {
            const int kv0 = t * 128;
            const int kv_valid = min(128, tokens - kv0);
            const int mode = (uniform && kv_valid == 128) ? 0 : (uniform ? 1 : 2);
            const uint32_t b = n_s & 1;
            const uint32_t s_col = t_lane + C::kColS + cq * 32;
            if (!a_loaded) {                                                // cold start: nothing was prefetched for this tile
                mbar_wait(&s_full[b], (n_s >> 1) & 1);
                tc_fence_after();
                tmem_ld16(s_col + b * 128, regA);
            }
            tmem_ld_wait16(regA);                                           // chunk A of S(n)
            uint32_t regB[16];
            tmem_ld16(s_col + b * 128 + 16, regB);                          // chunk B, consumed after the first 16 exponentials
            const bool ok_s = prefetch_next ? mbar_test(&s_full[b ^ 1], ((n_s + 1) >> 1) & 1) : true;
            const bool ok_p = n_s >= 2 ? mbar_test(&pv_done[b], ((n_s >> 1) - 1) & 1) : true;
            uint32_t packed[16];
            float lsum = 0.f;
            auto soft16 = [&](auto mode_c, const uint32_t* v, int half16) {
                constexpr int kMode = decltype(mode_c)::value;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float p0 = attn_exp2(fmaf(__uint_as_float(v[2 * i]), scale_log2e, -bound_log2e), i);
                    float p1 = attn_exp2(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2e, -bound_log2e), i);
                    if constexpr (kMode != 0) {
                        const int col = cq * 32 + half16 * 16 + 2 * i;
                        bool ok0 = col < kv_valid, ok1 = col + 1 < kv_valid;
                        if constexpr (kMode == 2) {
                            ok0 = ok0 && seg_s[kv0 + col] == my_seg;
                            ok1 = ok1 && seg_s[kv0 + col + 1] == my_seg;
                        }
                        p0 = ok0 ? p0 : 0.f;
                        p1 = ok1 ? p1 : 0.f;
                    }
                    packed[half16 * 8 + i] = Op16<OT>::pack(p0, p1);
                    lsum += p0 + p1;
                }
            };
            auto soft = [&](const uint32_t* v, int half16) {
                if (mode == 0) soft16(std::integral_constant<int, 0>{}, v, half16);
                else if (mode == 1) soft16(std::integral_constant<int, 1>{}, v, half16);
                else soft16(std::integral_constant<int, 2>{}, v, half16);
            };
            soft(regA, 0);
            if (pend) flush_p(b ^ 1);                                       // P(n-1): its store was issued a half tile ago
            tmem_ld_wait16(regB);                                           // chunk B: all of S(n) is in registers now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&s_free[b], leader);          // S(n+2) may overwrite the buffer
            if (prefetch_next) {                                            // chunk A of S(n+1)
                if (!ok_s) mbar_wait(&s_full[b ^ 1], ((n_s + 1) >> 1) & 1);
                tc_fence_after();
                tmem_ld16(s_col + (b ^ 1) * 128, regA);
            }
            a_loaded = prefetch_next;
            soft(regB, 1);
            if (!ok_p) mbar_wait(&pv_done[b], ((n_s >> 1) - 1) & 1);       // P V(n-2) has read this P buffer
            tc_fence_after();
            tmem_st16(t_lane + C::kColP + b * 64 + cq * 16, packed);
            pend = true;
            ++n_s;
            return lsum;
        };
        bool early_done = false;                                            // tile 0 of this item was already processed ...
        float l_early = 0.f;                                                // ... with this partial row sum
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            const bool uniform = uni_nx != 0;
            const float my_seg = seg_nx;
            fetch_meta(item + num_clusters);
            const int bh = item / q_pairs, qp = item - bh * q_pairs;
            const int qt = 2 * qp + (int)rank;                              // may be a phantom tile past the sequence: zeros in, clipped out
            const int sample = bh / heads, head = bh - sample * heads;
            const int nxt = item + num_clusters;
            const bool next_early = nxt < num_items && uni_nx != 0;         // the next item's first tile runs ahead of this item's epilogue
            if (!uniform) {                                                 // masked path: key segment ids of the sample in smem
                // (the pipeline never runs into a masked item: a_loaded and pend are false here)
                const float* segb = seg + (size_t)sample * tokens;
                for (int i = tid_sm; i < seg_pad; i += 512) seg_s[i] = i < tokens ? __ldg(segb + i) : 0.f;
                named_bar_sync(1, 512);
            }
            float l_run = early_done ? l_early : 0.f;
            for (int t = early_done ? 1 : 0; t < kv_tiles; ++t)
                l_run += do_tile(t, uniform, my_seg, t + 1 < kv_tiles || next_early);
            early_done = false;
            // The last P V of the item has not even been signalled yet: instead of draining the pipeline, the first tile of the NEXT
            // item -- its S was issued two tiles ago -- is processed now (which signals P of this item's last tile on its way).  Only
            // when the next item needs no segment table (a staged table would be overwritten under the slower warps of this item).
            if (next_early) { l_early = do_tile(0, true, seg_nx, kv_tiles > 1); early_done = true; }
            else if (pend) flush_p((n_s - 1) & 1);

            // ---- epilogue: O / l, zero padded queries (mask != 0), (128, DH) tile -> staging -> one TMA store ----
            l_part[cq * 128 + row] = l_run;
            if (store_pending) {                                            // the previous TMA store has read the staging tile
                if (tid_sm == 0) tma_store_wait_read();
                store_pending = false;
            }
            ATTN_TRACE(warp, 500);
            mbar_wait(o_full, n_item & 1);
            tc_fence_after();
            ATTN_TRACE(warp, 510);
            constexpr int OQ = C::kDHP / 4;                                 // output columns per thread: 20 or 24
            float o[OQ];
            tmem_ld16(t_lane + C::kColO + cq * OQ, reinterpret_cast<uint32_t*>(o));
            if constexpr (OQ == 20) tmem_ld4(t_lane + C::kColO + cq * OQ + 16, reinterpret_cast<uint32_t*>(o) + 16);
            else tmem_ld8(t_lane + C::kColO + cq * OQ + 16, reinterpret_cast<uint32_t*>(o) + 16);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(o_free, leader);
            named_bar_sync(1, 512);                                         // row sums exchanged; staging tile free
            const float l_tot = (l_part[row] + l_part[128 + row]) + (l_part[256 + row] + l_part[384 + row]);
            {
                const float inv = (my_seg != 0.f && l_tot > 0.f) ? 1.0f / l_tot : 0.f;
                const uint32_t dst = stage_sm + row * (DH * 2) + cq * (OQ * 2);
#pragma unroll
                for (int c = 0; c < OQ / 4; ++c) {
                    if (cq * OQ + c * 4 < DH)                               // skip the zero-pad columns 72..79
                        sts64(dst + c * 8, Op16<OT>::pack(o[c * 4] * inv, o[c * 4 + 1] * inv), Op16<OT>::pack(o[c * 4 + 2] * inv, o[c * 4 + 3] * inv));
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 512);
            if (tid_sm == 0) tma_store_4d(&map_o, stage_sm, 0, head, qt * 128, sample);   // rows >= tokens are clipped
            store_pending = true;
            ATTN_TRACE(warp, 530);
            ++n_item;
        }
        if (store_pending && tid_sm == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync();                                 // the pair must be done with this CTA's smem / TMEM / barriers
    if (warp == 1) { __syncwarp(); tmem_dealloc_2sm(tmem_base, 512); }
}

}  // namespace fitv2
