// Persistent, warp-specialised tcgen05 GEMM for the FiTv2 block projections.
//
//   C[M, N] = A[M, K] (16-bit, K-major) x W[N, K]^T (16-bit, K-major nn.Linear weight), fp32 accumulate in TMEM.
//
//   warp 0 : TMA producer  (cp.async.bulk.tensor, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1 : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-9 : epilogue.  Warp pairs (w, w+4) share a TMEM lane quarter and split the tile columns in two
//               halves.  tcgen05.ld gives every lane one accumulator ROW; a warp-private shared-memory slab
//               (XOR-swizzled, conflict-free) transposes it so that the global accesses are coalesced
//               (a warp instruction touches 4 full 128-byte lines instead of 32 partial ones).
//               TMEM accumulators are double-buffered: the epilogue of tile i overlaps the main loop of i+1.
//
// CL = 1: one CTA per 128 x BN tile, tcgen05.mma.cta_group::1.
// CL = 2: a cluster of two CTAs (one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2: each CTA
//         stages its own 128 rows of A and only HALF of the weight tile, the leader CTA issues the UMMAs, each
//         SM accumulates and post-processes its own 128 rows.  1-SM UMMA is bounded by shared-memory bandwidth
//         on Blackwell (operand reads + TMA writes share 128 B/clk/SM: measured 44% / 76% tensor-pipe
//         utilisation at BN = 144 / 256); the pair halves the B traffic per SM.
//
// Fused epilogues (reference lines they replace, paths relative to the reference repo):
//   EPI_QKV    bias + per-head LayerNorm(q,k) + interleaved-pair 2-D RoPE, scatter to Q / K / V^T
//              (fit/model/modules.py:166-174, fit/model/rope.py:107-111)
//   EPI_RESID  x += gate[sample] * (acc + bias)                      (modules.py:205,272 and :273)
//   EPI_SWIGLU hidden = silu(acc_gate + b_g) * (acc_up + b_x)        (timm SwiGLU; modules.py:251)
//   EPI_PLAIN  out = acc + bias                                      (generic nn.Linear; tests)
//   EPI_RESID_T  the same residual update computed TRANSPOSED: the weight rows are the M operand (128 output channels per
//              CTA, 256 per pair) and 256 token rows the N operand, so the tile is 256 columns wide and the main loop is
//              tensor-bound like gate/up instead of shared-memory bound (N = hidden_size only offers 128/144-wide tiles the
//              other way round).  TMEM lane = channel, so bias and gate are per-thread scalars.  The SMs never read the
//              residual: every warp stages gate * (acc + bias) for 16 token rows x 32 channels in shared memory and a TMA
//              reduce-add (cp.reduce.async.bulk.tensor .add, fp32) applies it to x in L2.  (A register version that loaded and
//              stored x itself was bound by the loads it could keep in flight: proj 102 us against 60.)
#pragma once
#include "common.cuh"
#include "tc2sm.cuh"

namespace fitv2 {

enum { EPI_QKV = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_PLAIN = 3, EPI_RESID_T = 4, EPI_QKV_GEN = 5, EPI_GELU = 6 };
// EPI_GELU: hidden = gelu_tanh(acc + bias) -> 16-bit, staged through the warp slabs like EPI_SWIGLU (timm Mlp fc1, modules.py:253).
// EPI_QKV_GEN: the QKV epilogue with a run-time q / k norm (none, LayerNorm with or without weight, RMSNorm with weight:
// fit/model/norms.py:35-50) for the configurations outside the FiTv2 default (affine-free LayerNorm, EPI_QKV).
__host__ __device__ constexpr bool epi_is_qkv(int epi) { return epi == EPI_QKV || epi == EPI_QKV_GEN; }
enum { QKNORM_NONE = 0, QKNORM_LN = 1, QKNORM_WLN = 2, QKNORM_RMS = 3 };

struct GemmEpi {
    const float* bias;        // [N] fp32 (layer slice)
    int tokens;               // tokens per sample row (m = sample * tokens + token)
    // EPI_RESID
    float* x;                 // [M, N] fp32 residual stream, updated in place
    const float* gate;        // [samples, gate_ld] fp32, already offset to this layer's gate chunk
    int gate_ld;
    // EPI_SWIGLU / EPI_PLAIN
    void* out16;              // [M, ld_out] 16-bit
    float* out32;             // [M, ld_out] fp32 (EPI_PLAIN when out16 == nullptr)
    int ld_out;
    // EPI_QKV
    void* q;                  // [samples, heads, tokens, DH]
    void* k;                  // [samples, heads, tokens, DH]
    void* vt;                 // [samples, heads, DH, tokens_v]   (V transposed: keys contiguous)
    const float* rope_cos;    // [DH/2, M]   (pair-major: lanes of a warp read consecutive token rows)
    const float* rope_sin;    // [DH/2, M]
    int heads;
    int tokens_v;
    // EPI_QKV_GEN
    int q_norm, k_norm;       // QKNORM_*
    const float* q_norm_w;    // [DH] (layer slice) or null
    const float* k_norm_w;
    // EPI_PLAIN
    const float* row_scale;   // [M] or null: out = (acc + bias) * row_scale[m]  (final layer: fit_model.py:230)
    int act_gelu;             // 1: out = gelu_tanh(acc + bias)  (timm Mlp fc1 + nn.GELU(approximate="tanh"), modules.py:253)
    // EPI_QKV_GEN
    int rope_v;               // 1: v is rotated like q / k (add_rel_pe_to_v, modules.py:171-172)
};

// nn.GELU(approximate="tanh"): 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3).  0.5 (1 + tanh(u)) = 1 / (1 + exp(-2u)), so
// gelu(x) = x * rcp(1 + exp2(w)) with w = -2 log2(e) u = x (c0 + c1 x^2): two MUFU operations (ex2, rcp) and five FP32 ones, like
// silu_mul (the IEEE-division form measured 249 us per fc1 launch at the headline shape: the epilogue, not the tensor pipe, was the
// bound).  Both ends saturate without NaN: exp2 -> inf gives rcp -> 0 (x * 0), exp2 -> 0 gives x.
__device__ __forceinline__ float gelu_tanh(float x) {
    constexpr float c0 = -2.0f * 1.4426950408889634f * 0.7978845608028654f;
    constexpr float c1 = c0 * 0.044715f;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + fast_exp2(x * fmaf(c1, x * x, c0))));
    return x * r;
}

constexpr int kGemmBM = 128;          // rows per CTA
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (GemmCfg::kThreads: the RESID epilogue may use 16 warps)
constexpr int kSmemBudget = 227 * 1024;

template <int BN, int EPI, int DH, int CL = 1> struct GemmCfg {
    static_assert(CL == 1 || CL == 2, "cluster size 1 or 2");
    static constexpr int kABytes = kGemmBM * kGemmBK * 2;
    static constexpr int kBRowsPerCta = BN / CL;                       // CL = 2: each CTA stages half of the weight tile
    static constexpr int kBBytes = kBRowsPerCta * kGemmBK * 2;
    // narrow tiles: one pipeline stage carries TWO 64-wide K blocks (8 UMMAs per barrier round trip); the UMMA issue /
    // commit / wait overhead per stage, not the tensor pipe, bounded BN <= 192 tiles (70 % tensor-pipe utilisation)
    static constexpr int kKSub = (kABytes + kBBytes) <= 28 * 1024 ? 2 : 1;
    // epilogue warps per TMEM lane quarter (each takes BN / kParts tile columns).  Four per quarter (16 epilogue warps)
    // only pay for the 256-wide residual tiles of the ragged-tiling experiments (proj 100 -> 79 us); at BN = 128 they
    // measured slightly slower than two (61.5 vs 57.3 us), so the production widths keep 8 epilogue warps.
    // QKV at head_dim 72: a 224-wide UMMA tile holds THREE heads (216 columns, one epilogue warp per quarter and head); the
    // last 8 columns are the first 8 of the next tile, computed and dropped (tiles advance by kTileN = 216).  The 144-wide
    // two-head tile is bound by shared-memory operand traffic (64 + 16384/BN bytes per tensor clock: 64 % tensor-pipe active).
    static constexpr bool kQkv3 = epi_is_qkv(EPI) && BN == 3 * DH + 8;
    static constexpr int kTileN = kQkv3 ? 3 * DH : BN;                  // output columns per tile
    static constexpr int kParts = (EPI == EPI_RESID && BN == 256) ? 4 : (kQkv3 ? 3 : 2);
    static constexpr int kThreads = 64 + kParts * 128;
    static constexpr int kStageBytes = (kABytes + kBBytes) * kKSub;    // per CTA
    static constexpr int kTxBytes = kStageBytes * CL;                  // credited to the (leader's) full barrier per stage
    static constexpr int kBarrierBytes = 1024;
    // epilogue staging: per warp 32 rows; QKV rows hold one head (16-bit) with a bank-conflict-free pitch,
    // the other epilogues use 128-byte XOR-swizzled rows.
    static constexpr int kEpiPitch = epi_is_qkv(EPI) ? ((DH * 2 / 4) % 8 == 4 ? DH * 2 : DH * 2 + 16) : 128;
    static constexpr int kEpiWarpBytes = 32 * kEpiPitch;
    static constexpr int kEpiBytes = (EPI == EPI_PLAIN ? 0 : 4 * kParts * kEpiWarpBytes);
    static constexpr int kStagesRaw = (kSmemBudget - kBarrierBytes - 1024 - kEpiBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarrierBytes + 1024;   // +1024 alignment slack
    static constexpr int kAccStride = 256;                                           // TMEM columns between the 2 accumulators
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128/256");
    static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment for SWIZZLE_128B");
    static_assert(kStages >= 3, "pipeline too shallow");
    static_assert(!epi_is_qkv(EPI) || DH * 64 <= kEpiWarpBytes, "V^T staging does not fit");
};

// ---- warp-private staging slab: 32 rows x 128 bytes, 16-byte chunk c of row r lives at chunk (c ^ (r & 7)) ----
__device__ __forceinline__ uint32_t slab_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

// Coalesced (row, chunk) owned by `lane` in iteration i when a slab row holds CPR 16-byte chunks.
template <int CPR> __device__ __forceinline__ void slab_task(int i, int lane, int& r, int& c) {
    const int task = i * 32 + lane;
    r = task / CPR;
    c = task % CPR;
}

// Work items of one cluster: item i of cluster c is tile c + i * clusters, n fastest (a wave shares few A tiles), or -- for the
// transposed residual GEMM -- row-tile groups fastest (a wave shares few token tiles).
struct TileWalk {
    int grp, num_groups, stride, n_tiles, bn, tile_n;
    int m_fast;                                                         // > 0: row-tile groups vary fastest (m_fast of them)
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ bool next(int& m_group, int& n0, int& bn_t) {
        if (grp >= num_groups) return false;
        if (m_fast) { m_group = grp % m_fast; n0 = (grp / m_fast) * tile_n; bn_t = bn; }
        else { m_group = grp / n_tiles; n0 = (grp % n_tiles) * tile_n; bn_t = bn; }
        grp += stride;
        return true;
    }
};

template <int BN, int EPI, typename OT, int DH, int CL>
__global__ void __launch_bounds__((GemmCfg<BN, EPI, DH, CL>::kThreads), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_x,   // EPI_RESID_T: fp32 residual stream (hidden_size, token rows), box 32 x 16
               int M, int N, int K, int b_row_offset, GemmEpi ep)
{
    using Cfg = GemmCfg<BN, EPI, DH, CL>;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    constexpr int KSUB = Cfg::kKSub;
    uint8_t* smem_b = smem + STAGES * Cfg::kABytes * KSUB;
    uint8_t* smem_epi = smem + STAGES * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
    const int n_tiles = EPI == EPI_RESID_T ? (N + Cfg::kTileN - 1) / Cfg::kTileN : N / Cfg::kTileN;   // RESID_T: N = token rows, ragged tail
    // work items are groups of CL vertically adjacent 128-row tiles; CTA `cta_rank` of the cluster owns row tile
    // group * CL + cta_rank (a phantom tile past the M tail computes on zero-filled rows and stores nothing)
    const int num_groups = ((m_tiles + CL - 1) / CL) * n_tiles;
    const int num_kb = (K + kGemmBK * Cfg::kKSub - 1) / (kGemmBK * Cfg::kKSub);       // pipeline stages per tile
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int group0 = blockIdx.x / CL, group_stride = gridDim.x / CL;
    const TileWalk walk0 = {group0, num_groups, group_stride, n_tiles, BN, Cfg::kTileN, EPI == EPI_RESID_T ? (m_tiles + CL - 1) / CL : 0};

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], Cfg::kParts * 128 * CL); }
        mbar_fence_init();
    }
    if (warp == 1) {
        if constexpr (CL == 2) tmem_alloc_2sm(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync(); else __syncthreads();   // the peer's barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                       // everything above overlapped the previous kernel's tail
    pdl_launch_dependents();

    // Producer and MMA loops are executed by the WHOLE warp (warp-uniform control flow lets the compiler keep the
    // descriptors / addresses in uniform registers, which the UTMALDG / UTCHMMA instructions consume directly);
    // one elected lane issues the asynchronous operations.
    if (warp == 0) {
        // ------------------------------ TMA producer (every CTA) ------------------------------
        int stage = 0; uint32_t phase = 0;
        TileWalk walk = walk0;
        walk.start();
        int m_group, n0, bn_t;
        while (walk.next(m_group, n0, bn_t)) {
            const int m_tile = m_group * CL + (int)cta_rank;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sa = smem_a + stage * Cfg::kABytes * KSUB;
                    uint8_t* sb = smem_b + stage * Cfg::kBBytes * KSUB;
                    // both CTAs' loads are credited to the leader's barrier, which expects the pair's bytes
                    if (CL == 1 || leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kTxBytes);
#pragma unroll
                    for (int j = 0; j < KSUB; ++j) {               // a K block past the end of K is zero-filled by TMA
                        const int kcol = (kb * KSUB + j) * kGemmBK;
                        if constexpr (EPI == EPI_RESID_T) {        // A = weight rows (layer offset), B = token rows
                            static_assert(EPI != EPI_RESID_T || CL == 2, "transposed residual GEMM is written for CTA pairs");
                            tma_load_2d_2sm(&tma_a, &full_bar[stage], sa + j * Cfg::kABytes, kcol, b_row_offset + m_tile * kGemmBM);
                            tma_load_2d_2sm(&tma_b, &full_bar[stage], sb + j * Cfg::kBBytes, kcol, n0 + (int)cta_rank * (bn_t / CL));
                        } else if constexpr (CL == 1) {
                            tma_load_2d(&tma_a, &full_bar[stage], sa + j * Cfg::kABytes, kcol, m_tile * kGemmBM);
                            tma_load_2d(&tma_b, &full_bar[stage], sb + j * Cfg::kBBytes, kcol, b_row_offset + n0);
                        } else {
                            tma_load_2d_2sm(&tma_a, &full_bar[stage], sa + j * Cfg::kABytes, kcol, m_tile * kGemmBM);
                            // a narrower (tail) tile still fetches the full box; the UMMA reads bn_t / CL rows of it
                            tma_load_2d_2sm(&tma_b, &full_bar[stage], sb + j * Cfg::kBBytes, kcol,
                                            b_row_offset + n0 + (int)cta_rank * (bn_t / CL));
                        }
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (CL == 1 || leader) {
            // ------------------------------ MMA issuer (leader CTA) ------------------------------
            constexpr uint32_t idesc = umma_idesc(Op16<OT>::kUmmaFormat, kGemmBM * CL, BN);
            const uint64_t da0 = umma_desc_kmajor(smem_u32(smem_a), 128);
            const uint64_t db0 = umma_desc_kmajor(smem_u32(smem_b), 128);
            int stage = 0; uint32_t phase = 0; int it = 0;
            TileWalk walk = walk0;
            walk.start();
            int m_group, n0, bn_t;
            for (; walk.next(m_group, n0, bn_t); ++it) {
                const uint32_t idesc_t = bn_t == BN ? idesc : umma_idesc(Op16<OT>::kUmmaFormat, kGemmBM * CL, bn_t);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);          // CL = 2: both CTAs' epilogues have drained it
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < KSUB; ++j) {
                            const uint64_t da = da0 + (uint64_t)((stage * KSUB + j) * (Cfg::kABytes >> 4));   // start-address field is addr >> 4
                            const uint64_t db = db0 + (uint64_t)((stage * KSUB + j) * (Cfg::kBBytes >> 4));
#pragma unroll
                            for (int kk = 0; kk < kGemmBK / 16; ++kk) {  // +32 bytes (>>4 = 2) per K=16 step inside the 128B atom
                                if constexpr (CL == 1) umma_ss(d_tmem, da + 2 * kk, db + 2 * kk, idesc_t, (kb | j | kk) != 0);
                                else umma_ss_2sm(d_tmem, da + 2 * kk, db + 2 * kk, idesc_t, (kb | j | kk) != 0);
                            }
                        }
                        // smem slot free (in both CTAs) once these MMAs retire; accumulator ready after the last k-block
                        if constexpr (CL == 1) {
                            umma_commit(&empty_bar[stage]);
                            if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                        } else {
                            umma_commit_2sm(&empty_bar[stage]);
                            if (kb == num_kb - 1) umma_commit_2sm(&tfull_bar[acc]);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ------------------------------ epilogue warps (every CTA: its own 128 rows) ------------------------------
        const int quarter = warp & 3;                                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                             // which BN / kParts column range of the tile this warp owns
        const uint32_t stg = smem_u32(smem_epi) + (warp - 2) * Cfg::kEpiWarpBytes;   // warp-private staging slab (shared-space address)
        // hand the accumulator back to the MMA issuer (which lives in the leader CTA)
        auto release_acc = [&](int acc) {
            tc_fence_before();
            if (CL == 1 || leader) mbar_arrive(&tempty_bar[acc]);
            else mbar_arrive_remote(&tempty_bar[acc], 0);
        };
        int it = 0;
        TileWalk walk = walk0;
        walk.start();
        int m_group, n0, bn_t;
        for (; walk.next(m_group, n0, bn_t); ++it) {
            const int m_tile = m_group * CL + (int)cta_rank, n_tile = n0 / Cfg::kTileN;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const uint32_t t_row = tmem_base + acc * Cfg::kAccStride + (uint32_t(quarter * 32) << 16);
            const int m_warp = m_tile * kGemmBM + quarter * 32;       // first row owned by this warp
            const int rows_valid = min(32, M - m_warp);               // <= 0 when the whole warp is past the M tail
            const int m = m_warp + lane;
            const bool row_ok = lane < rows_valid;
            const int s_first = m_warp / ep.tokens;
            const bool one_sample = rows_valid > 0 && (m_warp + rows_valid - 1) / ep.tokens == s_first;

            if constexpr (EPI == EPI_RESID) {
                constexpr int HALF = BN / Cfg::kParts;                // columns per warp
                constexpr int NFULL = HALF / 32, REM = HALF % 32;     // 32-column slabs + one remainder slab
                constexpr int NS = NFULL + (REM ? 1 : 0);
                constexpr int NPF = NS < 3 ? NS : 3;                  // slabs of x kept in flight (register ring)
                static_assert(REM == 0 || REM == 8 || REM == 16, "unsupported tile half width");
                const int col0 = n0 + half * HALF;
                // slabs of this warp inside the (possibly narrower, multiple-of-32 wide) tile
                const int ns_valid = REM == 0 ? max(0, min(NS, (bn_t - half * HALF) / 32)) : NS;
                float* xbase = ep.x + (size_t)m_warp * N + col0;
                // The residual does not depend on the accumulator: the x values of (up to) the whole tile half are
                // requested, in the coalesced mapping, BEFORE waiting for the MMAs of this tile, so the HBM latency
                // of the fp32 residual stream hides behind the main loop (it was the bound of the proj/fc2 GEMMs).
                float4 xr[NPF][8];
                auto load_x = [&](int slot, int sl) {                 // slot / sl are compile-time constants after unrolling
                    const int cpr = sl < NFULL ? 8 : REM / 4;         // 16-byte chunks per slab row
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < cpr) {
                            const int task = i * 32 + lane, r = task / cpr, c = task - r * cpr;
                            if (r < rows_valid && sl < ns_valid) xr[slot][i] = *reinterpret_cast<const float4*>(xbase + (size_t)r * N + sl * 32 + c * 4);
                        }
                    }
                };
#pragma unroll
                for (int sl = 0; sl < NPF; ++sl) load_x(sl, sl);
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                if (ns_valid == 0) release_acc(acc);
#pragma unroll
                for (int sl = 0; sl < NS; ++sl) {
                    if (sl < ns_valid) {
                    const int cpr = sl < NFULL ? 8 : REM / 4;
                    const int s0 = sl * 32;
                    uint32_t v[32];
                    if (cpr == 8) tmem_ld32(t_row + half * HALF + s0, v);
                    else if (cpr == 4) tmem_ld16(t_row + half * HALF + s0, v);
                    else tmem_ld8(t_row + half * HALF + s0, v);
                    tmem_ld_wait();
                    if (sl == ns_valid - 1) release_acc(acc);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < cpr) sts128(stg + slab_off(lane, c), make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
                    __syncwarp();
                    const int c0 = lane % cpr;                        // this lane's chunk column is the same in every iteration
                    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + s0 + c0 * 4));
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (one_sample) g = __ldg(reinterpret_cast<const float4*>(ep.gate + (size_t)s_first * ep.gate_ld + col0 + s0 + c0 * 4));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < cpr) {
                            const int task = i * 32 + lane, r = task / cpr, c = task - r * cpr;
                            if (r < rows_valid) {
                                if (!one_sample)
                                    g = __ldg(reinterpret_cast<const float4*>(ep.gate + (size_t)((m_warp + r) / ep.tokens) * ep.gate_ld + col0 + s0 + c * 4));
                                const uint4 au = lds128(stg + slab_off(r, c));
                                const float4 xv = xr[sl % NPF][i];
                                float4 o;
                                o.x = xv.x + g.x * (__uint_as_float(au.x) + b.x);
                                o.y = xv.y + g.y * (__uint_as_float(au.y) + b.y);
                                o.z = xv.z + g.z * (__uint_as_float(au.z) + b.z);
                                o.w = xv.w + g.w * (__uint_as_float(au.w) + b.w);
                                *reinterpret_cast<float4*>(xbase + (size_t)r * N + s0 + c * 4) = o;
                            }
                        }
                    }
                    __syncwarp();
                    if (sl + NPF < NS) load_x(sl % NPF, sl + NPF);    // refill the ring slot (only tiles wider than 3 slabs)
                    }
                }
            } else if constexpr (EPI == EPI_RESID_T) {
                // lane = output channel, columns = token rows.  M = channels (hidden_size), N = all token rows.
                constexpr int COLS = BN / Cfg::kParts;                // token rows per warp
                constexpr int CW = 16;                                // token rows per step (one tcgen05.ld.x16, one TMA box)
                constexpr int NCH = COLS / CW;
                static_assert(COLS % CW == 0 && Cfg::kEpiWarpBytes >= 2 * CW * 128, "transposed residual tile");
                const int ch0 = m_tile * kGemmBM + quarter * 32;       // first channel of this warp (a 128-byte line of x)
                const int ch = ch0 + lane;
                const bool ch_ok = ch < M;
                const int tok0 = n0 + half * COLS;
                const float b = ch_ok ? __ldg(ep.bias + ch) : 0.f;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    uint32_t v[CW];
                    tmem_ld16(t_row + half * COLS + c * CW, v);
                    tmem_ld_wait();
                    if (c == NCH - 1) release_acc(acc);
                    const int t_first = tok0 + c * CW;
                    const int s_a = t_first / ep.tokens;
                    const bool same = (t_first + CW - 1) / ep.tokens == s_a;
                    float g = 0.f;
                    if (same && ch_ok) g = __ldg(ep.gate + (size_t)s_a * ep.gate_ld + ch);
                    if (ch0 < M && t_first < N) {                       // warp-uniform; later steps of a tail tile are inactive too
                        // two staging buffers per warp: the reduce issued two steps ago must have read its buffer
                        if (lane == 0) tma_store_wait_read_n<1>();
                        __syncwarp();
                        const uint32_t buf = stg + (c & 1) * (CW * 128);
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            if (!same && ch_ok && t_first + j < N) g = __ldg(ep.gate + (size_t)((t_first + j) / ep.tokens) * ep.gate_ld + ch);
                            sts32(buf + j * 128 + lane * 4, __float_as_uint(g * (__uint_as_float(v[j]) + b)));
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        // rows past the last token and channels past hidden_size are clipped by the tensor map
                        if (lane == 0) tma_reduce_add_2d(&tma_x, buf, ch0, t_first);
                    }
                }
                if (lane == 0) tma_store_wait_read_n<0>();             // buffers are reused by the next tile
                __syncwarp();
            } else if constexpr (EPI == EPI_SWIGLU) {
                // tile columns: [0, BN/2) = gate rows of W, [BN/2, BN) = matching up rows (host packs W this way)
                constexpr int HALF = BN / 2;                          // outputs per tile
                constexpr int Q = HALF / 2;                           // outputs per warp: 64 -> one 128-byte slab row
                static_assert(Q == 64, "SwiGLU epilogue is written for 256-wide tiles");
                const float* bias_g = ep.bias + n0 + half * Q;
                const float* bias_u = ep.bias + n0 + HALF + half * Q;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < Q / 16; ++c) {
                    uint32_t g[16], u[16];
                    tmem_ld16(t_row + half * Q + c * 16, g);
                    tmem_ld16(t_row + HALF + half * Q + c * 16, u);
                    tmem_ld_wait();
                    if (c == Q / 16 - 1) release_acc(acc);
                    uint32_t packed[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bg = __ldg(reinterpret_cast<const float4*>(bias_g + c * 16 + j * 4));
                        const float4 bu = __ldg(reinterpret_cast<const float4*>(bias_u + c * 16 + j * 4));
                        packed[2 * j] = Op16<OT>::pack(silu_mul(__uint_as_float(g[4 * j]) + bg.x, __uint_as_float(u[4 * j]) + bu.x),
                                                       silu_mul(__uint_as_float(g[4 * j + 1]) + bg.y, __uint_as_float(u[4 * j + 1]) + bu.y));
                        packed[2 * j + 1] = Op16<OT>::pack(silu_mul(__uint_as_float(g[4 * j + 2]) + bg.z, __uint_as_float(u[4 * j + 2]) + bu.z),
                                                           silu_mul(__uint_as_float(g[4 * j + 3]) + bg.w, __uint_as_float(u[4 * j + 3]) + bu.w));
                    }
                    sts128(stg + slab_off(lane, 2 * c), make_uint4(packed[0], packed[1], packed[2], packed[3]));
                    sts128(stg + slab_off(lane, 2 * c + 1), make_uint4(packed[4], packed[5], packed[6], packed[7]));
                }
                __syncwarp();
                OT* obase = reinterpret_cast<OT*>(ep.out16) + (size_t)m_warp * ep.ld_out + n_tile * HALF + half * Q;
#pragma unroll
                for (int i = 0; i < 8; ++i) {                         // 4 rows x 128 contiguous bytes per warp instruction
                    int r, c; slab_task<8>(i, lane, r, c);
                    if (r < rows_valid)
                        *reinterpret_cast<uint4*>(obase + (size_t)r * ep.ld_out + c * 8) = lds128(stg + slab_off(r, c));
                }
                __syncwarp();
            } else if constexpr (EPI == EPI_GELU) {
                // timm Mlp fc1 + nn.GELU(approximate="tanh") (modules.py:253).  Each warp owns 128 tile columns and passes them through
                // its 4 KB slab in two 64-column halves: the staging and the 4 rows x 128 contiguous bytes store pattern of EPI_SWIGLU
                static_assert(BN == 256, "GELU epilogue is written for 256-wide tiles");
                constexpr int HALF = BN / 2;
                const float* bias = ep.bias + n0 + half * HALF;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t g[16];
                        tmem_ld16(t_row + half * HALF + pass * 64 + c * 16, g);
                        tmem_ld_wait();
                        if (pass == 1 && c == 3) release_acc(acc);
                        uint32_t packed[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + pass * 64 + c * 16 + j * 4));
                            packed[2 * j] = Op16<OT>::pack(gelu_tanh(__uint_as_float(g[4 * j]) + b.x), gelu_tanh(__uint_as_float(g[4 * j + 1]) + b.y));
                            packed[2 * j + 1] = Op16<OT>::pack(gelu_tanh(__uint_as_float(g[4 * j + 2]) + b.z), gelu_tanh(__uint_as_float(g[4 * j + 3]) + b.w));
                        }
                        sts128(stg + slab_off(lane, 2 * c), make_uint4(packed[0], packed[1], packed[2], packed[3]));
                        sts128(stg + slab_off(lane, 2 * c + 1), make_uint4(packed[4], packed[5], packed[6], packed[7]));
                    }
                    __syncwarp();
                    OT* obase = reinterpret_cast<OT*>(ep.out16) + (size_t)m_warp * ep.ld_out + n0 + half * HALF + pass * 64;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                     // 4 rows x 128 contiguous bytes per warp instruction
                        int r, c; slab_task<8>(i, lane, r, c);
                        if (r < rows_valid)
                            *reinterpret_cast<uint4*>(obase + (size_t)r * ep.ld_out + c * 8) = lds128(stg + slab_off(r, c));
                    }
                    __syncwarp();
                }
            } else if constexpr (EPI == EPI_PLAIN) {
                constexpr int HALF = BN / 2;
                constexpr int NG = HALF / 8;
                const float* bias = ep.bias + n0 + half * HALF;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < NG; ++c) {
                    uint32_t v[8];
                    tmem_ld8(t_row + half * HALF + c * 8, v);
                    tmem_ld_wait();
                    if (c == NG - 1) release_acc(acc);
                    if (row_ok) {
                        float o[8];
                        const float rs = ep.row_scale ? __ldg(ep.row_scale + m) : 1.0f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = (__uint_as_float(v[j]) + __ldg(bias + c * 8 + j)) * rs;
                        if (ep.act_gelu) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = gelu_tanh(o[j]);
                        }
                        if (ep.out16 != nullptr) {
                            OT* orow = reinterpret_cast<OT*>(ep.out16) + (size_t)m * ep.ld_out + n0 + half * HALF + c * 8;
                            *reinterpret_cast<uint4*>(orow) = make_uint4(Op16<OT>::pack(o[0], o[1]), Op16<OT>::pack(o[2], o[3]),
                                                                         Op16<OT>::pack(o[4], o[5]), Op16<OT>::pack(o[6], o[7]));
                        } else {
                            float* orow = ep.out32 + (size_t)m * ep.ld_out + n0 + half * HALF + c * 8;
                            *reinterpret_cast<float4*>(orow) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4*>(orow + 4) = make_float4(o[4], o[5], o[6], o[7]);
                        }
                    }
                }
            } else {   // EPI_QKV: every tile holds kParts whole heads, one per epilogue warp of a lane quarter; q / k / v is
                       // decided per head (a three-head tile may straddle the q|k or k|v boundary).
                static_assert(!epi_is_qkv(EPI) || BN == 2 * DH || Cfg::kQkv3, "QKV tile must be two or three heads wide");
                static_assert(DH % 8 == 0, "head_dim must be a multiple of 8");
                constexpr int CH = DH / 8;                            // 16-byte chunks per head row
                constexpr int PITCH = Cfg::kEpiPitch;
                const int ghead = n_tile * Cfg::kParts + half;
                const int kind = ghead / ep.heads;                    // 0 = q, 1 = k, 2 = v   (modules.py:166-167)
                const int head = ghead - kind * ep.heads;
                const float* bias = ep.bias + n0 + half * DH;
                const int token0 = m_warp - s_first * ep.tokens;      // token index of the warp's first row
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                float v[DH];
#pragma unroll
                for (int c = 0; c < DH / 8; ++c) tmem_ld8(t_row + half * DH + c * 8, reinterpret_cast<uint32_t*>(v) + c * 8);
                tmem_ld_wait();
                release_acc(acc);
#pragma unroll
                for (int c = 0; c < DH / 4; ++c) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c * 4));
                    v[c * 4] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
                }
                if (kind < 2) {
                    // LayerNorm over the head (no affine, eps 1e-6, biased variance): norms.py:41-42
                    float mean = 0.f, rstd = 1.f;
                    const int nmode = EPI == EPI_QKV ? QKNORM_LN : (kind == 0 ? ep.q_norm : ep.k_norm);
                    if (nmode == QKNORM_LN || nmode == QKNORM_WLN) {
                        float s4[4] = {0.f, 0.f, 0.f, 0.f};           // 4 independent chains instead of one 72-long one
#pragma unroll
                        for (int j = 0; j < DH; j += 4) { s4[0] += v[j]; s4[1] += v[j + 1]; s4[2] += v[j + 2]; s4[3] += v[j + 3]; }
                        mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / DH);
                        float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < DH; j += 4) {
                            const float d0 = v[j] - mean, d1 = v[j + 1] - mean, d2 = v[j + 2] - mean, d3 = v[j + 3] - mean;
                            q4[0] = fmaf(d0, d0, q4[0]); q4[1] = fmaf(d1, d1, q4[1]); q4[2] = fmaf(d2, d2, q4[2]); q4[3] = fmaf(d3, d3, q4[3]);
                        }
                        rstd = rsqrtf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.0f / DH) + 1e-6f);
                    } else if (nmode == QKNORM_RMS) {                  // norms.py:72-77: x * rsqrt(mean(x^2) + eps) * weight
                        float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < DH; j += 4) {
                            q4[0] = fmaf(v[j], v[j], q4[0]); q4[1] = fmaf(v[j + 1], v[j + 1], q4[1]);
                            q4[2] = fmaf(v[j + 2], v[j + 2], q4[2]); q4[3] = fmaf(v[j + 3], v[j + 3], q4[3]);
                        }
                        rstd = rsqrtf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.0f / DH) + 1e-6f);
                    }
                    if constexpr (EPI == EPI_QKV_GEN) {
                        const float* nw = kind == 0 ? ep.q_norm_w : ep.k_norm_w;
                        if (nmode != QKNORM_NONE) {
#pragma unroll
                            for (int j = 0; j < DH; ++j) v[j] = (v[j] - mean) * rstd;
                            mean = 0.f; rstd = 1.f;
                        }
                        if ((nmode == QKNORM_WLN || nmode == QKNORM_RMS) && nw != nullptr) {
#pragma unroll
                            for (int j = 0; j < DH; j += 4) {
                                const float4 w4 = __ldg(reinterpret_cast<const float4*>(nw + j));
                                v[j] *= w4.x; v[j + 1] *= w4.y; v[j + 2] *= w4.z; v[j + 3] *= w4.w;
                            }
                        }
                    }
                    const float* cs = ep.rope_cos + (row_ok ? m : 0);        // pair-major tables: coalesced across lanes
                    const float* sn = ep.rope_sin + (row_ok ? m : 0);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        uint32_t packed[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float cc = __ldg(cs + (size_t)(c * 4 + p) * M);
                            const float ss = __ldg(sn + (size_t)(c * 4 + p) * M);
                            const float x0 = (v[c * 8 + 2 * p] - mean) * rstd;
                            const float x1 = (v[c * 8 + 2 * p + 1] - mean) * rstd;
                            // q*cos + rotate_half(q)*sin with rotate_half: (x0,x1) -> (-x1,x0)
                            packed[p] = Op16<OT>::pack(x0 * cc - x1 * ss, x1 * cc + x0 * ss);
                        }
                        sts128(stg + lane * PITCH + c * 16, make_uint4(packed[0], packed[1], packed[2], packed[3]));
                    }
                    __syncwarp();
                    OT* dbase = reinterpret_cast<OT*>(kind == 0 ? ep.q : ep.k);
                    if (one_sample) {
                        // 32 token rows of one (sample, head) are contiguous in Q / K: flat 512-byte runs per instruction
                        OT* dst = dbase + (((size_t)s_first * ep.heads + head) * ep.tokens + token0) * DH;
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            const int idx = i * 32 + lane, r = idx / CH, c = idx - r * CH;
                            if (r < rows_valid)
                                *reinterpret_cast<uint4*>(dst + (size_t)idx * 8) = lds128(stg + r * PITCH + c * 16);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            const int idx = i * 32 + lane, r = idx / CH, c = idx - r * CH;
                            if (r < rows_valid) {
                                const int sm = (m_warp + r) / ep.tokens, tk = (m_warp + r) - sm * ep.tokens;
                                OT* dst = dbase + (((size_t)sm * ep.heads + head) * ep.tokens + tk) * DH;
                                *reinterpret_cast<uint4*>(dst + c * 8) = lds128(stg + r * PITCH + c * 16);
                            }
                        }
                    }
                    __syncwarp();
                } else {
                    if constexpr (EPI == EPI_QKV_GEN) {
                        if (ep.rope_v) {                                     // add_rel_pe_to_v: v * cos + rotate_half(v) * sin
                            const float* cs = ep.rope_cos + (row_ok ? m : 0);
                            const float* sn = ep.rope_sin + (row_ok ? m : 0);
#pragma unroll
                            for (int p = 0; p < DH / 2; ++p) {
                                const float cc = __ldg(cs + (size_t)p * M), ss = __ldg(sn + (size_t)p * M);
                                const float x0 = v[2 * p], x1 = v[2 * p + 1];
                                v[2 * p] = x0 * cc - x1 * ss;
                                v[2 * p + 1] = x1 * cc + x0 * ss;
                            }
                        }
                    }
                    // V^T[sample, head, d, token]
                    if (one_sample && rows_valid == 32 && (token0 & 7) == 0) {
                        // transpose through the slab: [d][32 tokens] 16-bit, then 16-byte (8-token) stores
#pragma unroll
                        for (int j = 0; j < DH; ++j) sts16(stg + (j * 32 + lane) * 2, Op16<OT>::bits(v[j]));
                        __syncwarp();
                        OT* dst = reinterpret_cast<OT*>(ep.vt) + ((size_t)s_first * ep.heads + head) * DH * ep.tokens_v + token0;
#pragma unroll
                        for (int i = 0; i < DH * 4 / 32; ++i) {
                            const int idx = i * 32 + lane, d = idx >> 2, part = idx & 3;
                            *reinterpret_cast<uint4*>(dst + (size_t)d * ep.tokens_v + part * 8) = lds128(stg + d * 64 + part * 16);
                        }
                        __syncwarp();
                    } else if (row_ok) {
                        const int sm = m / ep.tokens, tk = m - sm * ep.tokens;
                        OT* dst = reinterpret_cast<OT*>(ep.vt) + ((size_t)sm * ep.heads + head) * DH * ep.tokens_v + tk;
#pragma unroll
                        for (int j = 0; j < DH; ++j) dst[(size_t)j * ep.tokens_v] = Op16<OT>::from(v[j]);
                    }
                }
            }
        }
    }

    if constexpr (EPI == EPI_RESID_T) {
        if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // reduce-adds performed before exit
    }
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync(); else __syncthreads();   // the pair must be done with this CTA's smem / TMEM / barriers
    if (warp == 1) {
        __syncwarp();
        if constexpr (CL == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace fitv2
