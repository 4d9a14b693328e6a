// C-ABI of the B200-native FiTv2 hot path (see include/fitv2_b200.h).  Host orchestration only:
// workspace carving, TMA descriptors, kernel selection and the per-NFE launch sequence that
// replaces fit/model/fit_model.py:189-233 (FiT.forward) of the reference.
#include "../../include/fitv2_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "attention_general.cuh"
#include "attention_ws.cuh"
#include "attention_tm.cuh"
#include "pointwise.cuh"
#include "cond_tc.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <unordered_map>
#include <vector>

using namespace fitv2;

namespace {

// scratch for the split-K partial sums of the conditioning linears
constexpr size_t kCondPartialBytes = 24u << 20;

// Thread-block cluster size of the production GEMMs (weight-tile TMA multicast across the cluster).
constexpr int kGemmCluster = 2;

thread_local std::string g_last_error;

// Programmatic dependent launch for every kernel of the chain (see pdl_wait() in common.cuh).  Per handle (option "pdl");
// the handle-less elementwise entry points always use it.
thread_local bool g_pdl = true;

// L2 residency window of the fp32 residual stream (set per forward): kernels launched while it is active carry an
// access-policy window that marks `hit_ratio` of its lines persisting in L2 and everything else streaming, so that the
// residual written by one kernel is still in L2 when the next one reads it (it is touched six times per block).
struct L2Window { void* base = nullptr; size_t bytes = 0; float hit_ratio = 0.f; };
thread_local L2Window g_l2_window;

template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[3];
    unsigned n = 0;
    if (g_l2_window.base) {
        attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[n].val.accessPolicyWindow.base_ptr = g_l2_window.base;
        attr[n].val.accessPolicyWindow.num_bytes = g_l2_window.bytes;
        attr[n].val.accessPolicyWindow.hitRatio = g_l2_window.hit_ratio;
        attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        ++n;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (g_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(FITV2_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D K-major operand map: tensor (rows, cols) 16-bit with row pitch ld elements; box = 64 cols x box_rows,
// 128B swizzle, out-of-bounds reads return zeros (M / K tails).
int make_map(CUtensorMap* map, const void* ptr, int operand_dtype, uint64_t rows, uint64_t cols, uint64_t ld,
             uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, operand_dtype == FITV2_OPERAND_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u ptr=%p", (int)r,
                    (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, ptr);
    return FITV2_OK;
}

// 2-D K-major fp32 (tf32 operand) map: box = 32 cols (128 bytes) x box_rows, 128B swizzle, zero fill out of bounds.
int make_map_f32(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled(fp32) failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u ptr=%p", (int)r,
                    (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, ptr);
    return FITV2_OK;
}

// 3-D map (d0 contiguous, d1, d2) with a (box0, box1, 1) box; swizzle_bytes = 128 / 64 / 32 must equal box0 * 2.
int make_map3(CUtensorMap* map, const void* ptr, int operand_dtype, uint64_t d0, uint64_t d1, uint64_t d2,
              uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(map, operand_dtype == FITV2_OPERAND_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d) dims=%llu,%llu,%llu box=%u,%u ptr=%p", (int)r,
                    (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box0, box1, ptr);
    return FITV2_OK;
}

// 4-D output map of the attention kernel: (head_dim, heads, tokens, samples) over the (M, heads*head_dim) 16-bit matrix,
// box = one (128 tokens x head_dim) tile of one head, no swizzle (dense staging rows).
int make_map_attn_out(CUtensorMap* map, const void* ptr, int operand_dtype, uint64_t dh, uint64_t heads, uint64_t tokens,
                      uint64_t samples) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[4] = {dh, heads, tokens, samples};
    cuuint64_t strides[3] = {dh * 2, heads * dh * 2, tokens * heads * dh * 2};
    cuuint32_t box[4] = {(cuuint32_t)dh, 1, 128, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, operand_dtype == FITV2_OPERAND_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled(attention out) failed (%d)", (int)r);
    return FITV2_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
    int rows = 0, tokens = 0, tokens_v = 0;
    int n_spans = 0;                                   // (offset, bytes) of every buffer, in allocation order (fitv2_debug_layout)
    size_t span_off[40], span_bytes[40];
    size_t x_res, h, ao, q, k, vt, hidden, te, t0, c, sc, sc_split, lmid, lmid_split, gmod, mod, fmod, rope_cos, rope_sin, seg_uniform, cpart, xt, ot, wfin16,
           sg_g, sg_s, sg_m, total;
};

}  // namespace

enum { PC_COND = 0, PC_LNMOD, PC_QKV, PC_ATTN, PC_PROJ, PC_GATEUP, PC_FC2, PC_MISC, PC_COUNT };
struct ProfEvt { cudaEvent_t a, b; int cls; };

// Per-handle tuning switches (fitv2_set_option).  They used to be FITV2_* environment variables latched in function-local
// statics; now every handle carries its own copy and the Python layer forwards the environment at handle creation.
struct Options {
    int pdl = 1;                 // programmatic dependent launch on every kernel of the chain
    int attn = 0;                // 0 auto, 1 attention_tm, 2 attention_ws, 3 attention_general (online max)
    int attn_early = 1;          // attention_tm: softmax of the next item's first sub-tile ahead of the current epilogue
    int ln_threads = 64;         // threads per CTA of the LayerNorm + modulate kernel (one row per warp)
    int ln_wide_single = 0;      // hidden 2304: one warp per row instead of a warp pair
    int bn_resid = 0;            // tile width of the proj / fc2 GEMMs in the normal orientation (0 = cost model)
    int qkv_heads = 3;           // heads per QKV tile at head_dim 72 (3 -> 224-wide tile, 2 -> 144)
    int resid_t = -1;            // transposed residual GEMM + TMA reduce-add: -1 auto (fc2 only, by cost model), 0 off, 1 proj and fc2, 2 fc2
    int bn_resid_t = 0;          // token rows per transposed tile (0 = cost model, 224, 256)
    int cond = 0;                // conditioning linears: 0 tensor pipe (tf32) where the shapes tile, 1 fp32 FMA kernels
    int l2_persist_mb = 0;       // persisting-L2 window on the fp32 residual stream (0 = off)
    int ws_guard = 0;            // bytes of guard padding behind every workspace buffer (bounds test)
    int final_tc = 1;            // final layer: 1 = LayerNorm+modulate kernel + skinny tcgen05 GEMM, 0 = fused fp32 SIMT kernel
    int gelu_epi = 0;            // GELU Mlp fc1: 0 = slab-staged epilogue (EPI_GELU), 1 = plain epilogue with act_gelu (16-byte row stores)
    int verbose = 0;
};

struct AttnMaps {                // TMA descriptors of one attention launch, cached per (pointers, rows, tokens)
    const void *q = nullptr, *k = nullptr, *vt = nullptr, *out = nullptr;
    int rows = 0, tokens = 0;
    CUtensorMap mq, mqt, mk, mkt, mv, mo;
};

struct fitv2_handle {
    fitv2_config cfg;
    Options opt;
    int device = 0;
    const void* w[FITV2_W_COUNT];
    int64_t w_numel[FITV2_W_COUNT];
    uint8_t* ws = nullptr;
    int64_t ws_bytes = 0;
    Layout lay;
    bool maps_valid = false;
    CUtensorMap map_h, map_ao, map_hidden;              // activations (A operands)
    CUtensorMap map_wqkv, map_wproj, map_wgu, map_wfc2; // stacked weights (B operands)
    // conditioning on the tensor pipe (cond_tc.cuh): fp32 / tf32 maps of the split activations and of the adaLN weights
    bool cond_tc = false;
    int cond_bn_up = 0;
    CUtensorMap map_sc_split, map_lmid_split, map_wglobal, map_wfinal, map_wlora_a, map_wlora_b, map_wnormal;
    int bn_proj = 0, bn_fc2 = 0;
    bool proj_t = false, fc2_t = false;                  // proj / fc2 run as transposed 256-token-wide tiles (EPI_RESID_T)
    CUtensorMap map_wproj_t, map_wfc2_t;                 // the same weights with 128-row boxes (M operand of EPI_RESID_T)
    CUtensorMap map_ao_t, map_hidden_t;                  // activations as the N operand of EPI_RESID_T: bn_resid_t / 2 token rows per box
    int bn_resid_t = 256;                                // token rows per transposed tile (256 or 224, whichever needs fewer tensor clocks)
    CUtensorMap map_x;                                   // fp32 residual stream, 32-channel x 16-token boxes (TMA reduce-add target)
    bool qkv3 = false;                                   // QKV GEMM uses the three-head 224-wide tile (head_dim 72)
    bool qkv_gen = false;                                // q / k norm other than affine-free LayerNorm: EPI_QKV_GEN + attention_general
    bool final_tc = false;                               // final linear on the tensor pipe (16-bit copy of the weight in the workspace)
    bool wfin16_valid = false;                           // ... which is refreshed by the first forward after a (re)bind / new workspace / new layout
    CUtensorMap map_hfinal, map_wfinal_lin;
    AttnMaps attn_maps;
    int num_sms = 148;
    const float* online_fh = nullptr;                    // per-row RoPE frequencies (rows, head_dim/4), online_rope mode
    const float* online_fw = nullptr;
    int online_rows = 0;
    int64_t l2_persist_bytes = 0, l2_window_max = 0;     // persisting-L2 carve-out used for the residual stream (0 = off)
    int64_t launches = 0;
    // kernels whose MaxDynamicSharedMemorySize attribute has been raised ON THIS HANDLE'S DEVICE (function attributes are per
    // device; a process-wide latch skipped the call on a second GPU)
    std::unordered_map<const void*, int> smem_configured;
    // sticky device-side error word in pinned, mapped host memory (bit 0: class label outside the embedding table)
    int* err_host = nullptr;
    int* err_dev = nullptr;
    // optional per-kernel-class CUDA-event timing (fitv2_profile_*)
    uint32_t prof_mask = 0;
    std::vector<ProfEvt> prof_pool;
    size_t prof_used = 0;
    int prof_open = -1;
};

namespace {

int out_channels(const fitv2_config& c) { return c.out_channels > 0 ? c.out_channels : c.token_channels; }
bool norm_has_weight(int mode) { return mode == FITV2_NORM_WLAYERNORM || mode == FITV2_NORM_RMSNORM; }

// Element count of a weight slot for this configuration; 0 = the slot is not used (and need not be bound).
int64_t expected_numel(const fitv2_config& c, int slot) {
    const int64_t D = c.hidden_size, L = c.depth, Hm = c.mlp_hidden, lora = c.lora_dim, C = c.token_channels, Co = out_channels(c);
    const bool lora_mode = c.adaln_type == FITV2_ADALN_LORA, normal_mode = c.adaln_type == FITV2_ADALN_NORMAL, sg = c.adaln_type == FITV2_ADALN_SWIGLU;
    const int64_t Hs = (D / 4) * 3, Hf = D / 2;                       // modules.py:266,285
    switch (slot) {
        case FITV2_W_X_EMBED_W: return D * C;
        case FITV2_W_X_EMBED_B: return D;
        case FITV2_W_T_MLP0_W: return D * 256;
        case FITV2_W_T_MLP0_B: return D;
        case FITV2_W_T_MLP2_W: return D * D;
        case FITV2_W_T_MLP2_B: return D;
        case FITV2_W_Y_TABLE: return (int64_t)c.num_embeddings * D;
        case FITV2_W_GLOBAL_ADALN_W: return lora_mode ? 6 * D * D : 0;
        case FITV2_W_GLOBAL_ADALN_B: return lora_mode ? 6 * D : 0;
        case FITV2_W_LORA_A_W: return lora_mode ? L * lora * D : 0;
        case FITV2_W_LORA_A_B: return lora_mode ? L * lora : 0;
        case FITV2_W_LORA_B_W: return lora_mode ? L * 6 * D * lora : 0;
        case FITV2_W_LORA_B_B: return lora_mode ? L * 6 * D : 0;
        case FITV2_W_FINAL_ADALN_W: return sg ? 0 : 2 * D * D;
        case FITV2_W_FINAL_ADALN_B: return sg ? 0 : 2 * D;
        case FITV2_W_FINAL_LINEAR_W: return Co * D;
        case FITV2_W_FINAL_LINEAR_B: return Co;
        case FITV2_W_QKV_W: return L * 3 * D * D;
        case FITV2_W_QKV_B: return L * 3 * D;
        case FITV2_W_PROJ_W: return L * D * D;
        case FITV2_W_PROJ_B: return L * D;
        case FITV2_W_GATEUP_W: return L * (c.mlp_type == FITV2_MLP_GELU ? 1 : 2) * Hm * D;    // GELU Mlp: fc1 alone
        case FITV2_W_GATEUP_B: return L * (c.mlp_type == FITV2_MLP_GELU ? 1 : 2) * Hm;
        case FITV2_W_FC2_W: return L * D * Hm;
        case FITV2_W_FC2_B: return L * D;
        case FITV2_W_ROPE_FREQS_H: return c.head_dim / 4;
        case FITV2_W_ROPE_FREQS_W: return c.head_dim / 4;
        case FITV2_W_NORMAL_ADALN_W: return normal_mode ? L * 6 * D * D : 0;
        case FITV2_W_NORMAL_ADALN_B: return normal_mode ? L * 6 * D : 0;
        case FITV2_W_SG_G_W: case FITV2_W_SG_X_W: return sg ? L * Hs * D : 0;
        case FITV2_W_SG_G_B: case FITV2_W_SG_X_B: return sg ? L * Hs : 0;
        case FITV2_W_SG_FC2_W: return sg ? L * 6 * D * Hs : 0;
        case FITV2_W_SG_FC2_B: return sg ? L * 6 * D : 0;
        case FITV2_W_FSG_G_W: case FITV2_W_FSG_X_W: return sg ? Hf * D : 0;
        case FITV2_W_FSG_G_B: case FITV2_W_FSG_X_B: return sg ? Hf : 0;
        case FITV2_W_FSG_FC2_W: return sg ? 2 * D * Hf : 0;
        case FITV2_W_FSG_FC2_B: return sg ? 2 * D : 0;
        case FITV2_W_NORM1_W: return norm_has_weight(c.block_norm) ? L * D : 0;
        case FITV2_W_NORM2_W: return norm_has_weight(c.block_norm) ? L * D : 0;
        case FITV2_W_NORM_FINAL_W: return norm_has_weight(c.block_norm) ? D : 0;
        case FITV2_W_Q_NORM_W: return norm_has_weight(c.q_norm) ? L * c.head_dim : 0;
        case FITV2_W_K_NORM_W: return norm_has_weight(c.k_norm) ? L * c.head_dim : 0;
    }
    return -1;
}

Layout make_layout(const fitv2_config& c, int rows, int tokens, size_t guard = 0) {
    Layout l;
    l.rows = rows; l.tokens = tokens; l.tokens_v = (tokens + 7) / 8 * 8;
    const size_t M = (size_t)rows * tokens, D = c.hidden_size, Hm = c.mlp_hidden, L = c.depth;
    const size_t lora = c.adaln_type == FITV2_ADALN_LORA ? c.lora_dim : 0;
    size_t off = 0;
    // `guard` bytes of untouched padding behind every buffer (option "ws_guard"): the bounds test fills the workspace with a
    // canary, runs a forward and checks that the padding still holds it
    auto take = [&](size_t bytes) {
        size_t o = off;
        if (l.n_spans < 40) { l.span_off[l.n_spans] = o; l.span_bytes[l.n_spans] = bytes; ++l.n_spans; }
        off = align_up(off + bytes + guard, 1024);
        return o;
    };
    l.x_res = take(M * D * 4);
    l.h = take(M * D * 2);
    l.ao = take(M * D * 2);
    l.q = take(M * D * 2);
    l.k = take(M * D * 2);
    l.vt = take((size_t)rows * D * l.tokens_v * 2);
    l.hidden = take(M * Hm * 2);
    l.te = take((size_t)rows * 256 * 4);
    l.t0 = take((size_t)rows * D * 4);
    l.c = take((size_t)rows * D * 4);
    l.sc = take((size_t)rows * D * 4);
    l.lmid = take((size_t)rows * L * lora * 4);
    // hi / lo stacked tf32 operands of the tensor-pipe conditioning linears (cond_tc.cuh): 128 rows per 64 logical rows
    const size_t split_rows = (size_t)((rows + kCondRows - 1) / kCondRows) * 128;
    l.sc_split = take(split_rows * D * 4);
    l.lmid_split = take(split_rows * L * lora * 4);
    l.gmod = take((size_t)rows * 6 * D * 4);
    l.mod = take(L * (size_t)rows * 6 * D * 4);
    l.fmod = take((size_t)rows * 2 * D * 4);
    l.rope_cos = take(M * (c.head_dim / 2) * 4);
    l.rope_sin = take(M * (c.head_dim / 2) * 4);
    l.seg_uniform = take((size_t)rows * 4);
    l.cpart = take(kCondPartialBytes);
    // channels-first callers (use_sit = False: (B, C, N) tensors): token-major copies of the input / output
    l.xt = take(c.channels_first ? M * c.token_channels * 4 : 0);
    l.ot = take(c.channels_first ? M * out_channels(c) * 4 : 0);
    l.wfin16 = take((size_t)out_channels(c) * D * 2);                   // fp16 copy of final_layer.linear.weight (tensor-pipe final layer)
    // adaln_type 'swiglu': fc1_g(c), silu(fc1_g(c)) and the gated hidden of all blocks, (L, rows, (D/4)*3) fp32 each
    const size_t sg_bytes = c.adaln_type == FITV2_ADALN_SWIGLU ? L * (size_t)rows * ((D / 4) * 3) * 4 : 0;
    l.sg_g = take(sg_bytes);
    l.sg_s = take(sg_bytes);
    l.sg_m = take(sg_bytes);
    l.total = off;
    return l;
}

// Event-bracket the launches of one kernel class on the launching stream (only when enabled).
inline void prof_begin(fitv2_handle* h, int cls, cudaStream_t st) {
    if (!(h->prof_mask & (1u << cls)) || h->prof_used >= h->prof_pool.size()) return;
    ProfEvt& e = h->prof_pool[h->prof_used];
    e.cls = cls;
    cudaEventRecord(e.a, st);
    h->prof_open = (int)h->prof_used;
}
inline void prof_end(fitv2_handle* h, cudaStream_t st) {
    if (h->prof_open < 0) return;
    cudaEventRecord(h->prof_pool[h->prof_open].b, st);
    h->prof_used++;
    h->prof_open = -1;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: remember per handle (= per device) what was set.
template <typename K>
int ensure_smem(fitv2_handle* h, K kern, int bytes) {
    int& have = h->smem_configured[reinterpret_cast<const void*>(kern)];
    if (have < bytes) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have = bytes;
    }
    return FITV2_OK;
}

template <int BN, int EPI, typename OT, int DH, int CL = kGemmCluster>
int launch_gemm_t(fitv2_handle* h, const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, int b_row_off,
                  const GemmEpi& ep, cudaStream_t st) {
    using Cfg = GemmCfg<BN, EPI, DH, CL>;
    auto kern = gemm_tc_kernel<BN, EPI, OT, DH, CL>;
    int rc = ensure_smem(h, kern, Cfg::kSmemBytes);
    if (rc) return rc;
    constexpr int TN = Cfg::kTileN;
    // EPI_RESID_T: N = token rows, the tail tile is zero-filled / clipped; every other epilogue needs whole tiles
    if (EPI != EPI_RESID_T && N % TN != 0)
        return fail(FITV2_E_INVALID, "GEMM N=%d is not a multiple of the tile width %d", N, TN);
    const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
    const int groups = ((m_tiles + CL - 1) / CL) * ((N + TN - 1) / TN);
    const int max_clusters = h->num_sms / CL;
    const int grid = (groups < max_clusters ? groups : max_clusters) * CL;
    CUDA_TRY(launch_k(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, CL, ma, mb, h->map_x, M, N, K, b_row_off, ep));
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return FITV2_OK;
}

template <int EPI, typename OT, int CL = kGemmCluster>
int launch_gemm_bn(fitv2_handle* h, int bn, const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K,
                   int b_row_off, const GemmEpi& ep, cudaStream_t st) {
    switch (bn) {
        case 128: return launch_gemm_t<128, EPI, OT, 0, CL>(h, ma, mb, M, N, K, b_row_off, ep, st);
        case 144: return launch_gemm_t<144, EPI, OT, 0, CL>(h, ma, mb, M, N, K, b_row_off, ep, st);
        case 192: return launch_gemm_t<192, EPI, OT, 0, CL>(h, ma, mb, M, N, K, b_row_off, ep, st);
        case 256: return launch_gemm_t<256, EPI, OT, 0, CL>(h, ma, mb, M, N, K, b_row_off, ep, st);
    }
    return fail(FITV2_E_INVALID, "unsupported GEMM tile width %d", bn);
}

// Tile width for the N = hidden_size projections among the uniform widths that divide N: minimise (waves x per-tile cost).
// Cost of one 64-deep K block of a (128 rows per CTA) x bn tile in SM clocks: the tensor pipe needs 2*bn, the shared memory
// pipe (TMA writes + UMMA operand reads share 128 B/clk) needs (16384 + 64*bn) * 2 / 128 = 256 + bn, so tiles narrower than
// 256 columns are shared-memory bound (measured: 72 % tensor-pipe ceiling at bn = 144).  A ragged 256 + tail tiling was
// measured slower for these two GEMMs (they are bound by the residual epilogue, profiles/README.md) and has been removed.
int pick_bn(int M, int N, int num_sms, bool prefer_aligned) {
    const int cands[4] = {256, 192, 144, 128};
    int best = 0; long best_cost = 0;
    for (int bn : cands) {
        if (N % bn) continue;
        const long m_tiles = (M + kGemmBM - 1) / kGemmBM;
        const long tiles = ((m_tiles + kGemmCluster - 1) / kGemmCluster) * kGemmCluster * (N / bn);
        const long waves = (tiles + num_sms - 1) / num_sms;
        long cost = waves * (bn + 16);
        if (prefer_aligned && (bn / 2) % 32 != 0) cost += cost / 4;    // tile halves not aligned to 128-byte lines of the fp32 residual
        if (!best || cost < best_cost) { best = bn; best_cost = cost; }
    }
    return best;
}

template <typename OT>
int launch_attention(fitv2_handle* h, const void* q, const void* k, const void* vt, const float* seg,
                     const int* seg_uniform, void* out, int rows, int tokens, int tokens_v, cudaStream_t st,
                     float* dbg_s = nullptr, float* dbg_o = nullptr) {
    const fitv2_config& c = h->cfg;
    const float scale_log2e = (1.0f / sqrtf((float)c.head_dim)) * 1.4426950408889634f;
    // logit bound from the affine-free QK-LayerNorm (attention_ws.cuh): |q.k|/sqrt(dh) <= mag^2 * sqrt(dh); 2% margin.
    // fp16 operands: shift by 8 binades so that P stays in the normal fp16 range (cancels in O / l).
    const float bound_log2e = c.rope_magnitude * c.rope_magnitude * sqrtf((float)c.head_dim) * 1.02f * 1.4426950408889634f -
                              (c.operand_dtype == FITV2_OPERAND_FP16 ? 8.0f : 0.0f);
    // TMA maps over Q / K (dh, tokens, rows*heads), V^T (tokens_v, dh, rows*heads) and the output; out-of-bounds = zero fill.
    // Encoded once per (buffers, rows, tokens): a forward launches this `depth` times with the same workspace buffers.
    AttnMaps& am = h->attn_maps;
    if (am.q != q || am.k != k || am.vt != vt || am.out != out || am.rows != rows || am.tokens != tokens) {
        const uint64_t DHu = c.head_dim, BH = (uint64_t)rows * c.num_heads;
        const int dhp = (c.head_dim + 15) / 16 * 16, tail = dhp - 64;
        int rc;
        am.q = nullptr;
        if ((rc = make_map3(&am.mq, q, c.operand_dtype, DHu, tokens, BH, DHu * 2, (uint64_t)tokens * DHu * 2, 64, 128, 128))) return rc;
        if ((rc = make_map3(&am.mqt, q, c.operand_dtype, DHu, tokens, BH, DHu * 2, (uint64_t)tokens * DHu * 2, tail, 128, tail * 2))) return rc;
        if ((rc = make_map3(&am.mk, k, c.operand_dtype, DHu, tokens, BH, DHu * 2, (uint64_t)tokens * DHu * 2, 64, 128, 128))) return rc;
        if ((rc = make_map3(&am.mkt, k, c.operand_dtype, DHu, tokens, BH, DHu * 2, (uint64_t)tokens * DHu * 2, tail, 128, tail * 2))) return rc;
        // V^T rows are tokens_v (a multiple of 8) long in memory, but only `tokens` keys exist: the map ends at `tokens`, so the pad
        // columns (never written by the QKV epilogue) are zero-filled by TMA instead of read (0 * NaN garbage = NaN in P V)
        if ((rc = make_map3(&am.mv, vt, c.operand_dtype, tokens, DHu, BH, (uint64_t)tokens_v * 2, DHu * tokens_v * 2, 64, dhp, 128))) return rc;
        if ((rc = make_map_attn_out(&am.mo, out, c.operand_dtype, DHu, c.num_heads, tokens, rows))) return rc;
        am.q = q; am.k = k; am.vt = vt; am.out = out; am.rows = rows; am.tokens = tokens;
    }
    // Kernel choice.  Affine-free QK-LayerNorm (the FiTv2 family): bound-based softmax, P in tensor memory up to 256 tokens at
    // head_dim 72 (attention_tm.cuh), the shared-memory-P pipeline beyond (attention_ws.cuh).  Any other q / k norm has
    // unbounded logits: online-max kernel (attention_general.cuh), which also serves the debug taps.
    const bool bounded = c.q_norm == FITV2_NORM_LAYERNORM && c.k_norm == FITV2_NORM_LAYERNORM;
    int mode = h->opt.attn;
    if (!bounded || dbg_s || dbg_o) mode = 3;
    if (mode == 0) mode = (tokens <= 256 && c.head_dim == 72) ? 1 : 2;
    if ((mode == 1 || mode == 4) && c.head_dim != 72) mode = 2;
    if (mode == 3) {
        const int num_items = ((tokens + 127) / 128) * c.num_heads * rows;      // (query tile, head, sample) work items
        const int grid = num_items < 2 * h->num_sms ? num_items : 2 * h->num_sms;   // persistent, two CTAs per SM
        int rc;
        if (c.head_dim == 72) {
            auto kern = attention_general_kernel<OT, 72>;
            if ((rc = ensure_smem(h, kern, AttnGenCfg<72>::kSmemBytes))) return rc;
            CUDA_TRY(launch_k(kern, grid, dim3(AttnGenCfg<72>::kThreads), AttnGenCfg<72>::kSmemBytes, st, 1, am.mq, am.mqt, am.mk, am.mkt, am.mv,
                              seg, seg_uniform, (OT*)out, c.num_heads, tokens, num_items, scale_log2e, dbg_s, dbg_o));
        } else {
            auto kern = attention_general_kernel<OT, 96>;
            if ((rc = ensure_smem(h, kern, AttnGenCfg<96>::kSmemBytes))) return rc;
            CUDA_TRY(launch_k(kern, grid, dim3(AttnGenCfg<96>::kThreads), AttnGenCfg<96>::kSmemBytes, st, 1, am.mq, am.mqt, am.mk, am.mkt, am.mv,
                              seg, seg_uniform, (OT*)out, c.num_heads, tokens, num_items, scale_log2e, dbg_s, dbg_o));
        }
        CUDA_TRY(cudaGetLastError());
        h->launches++;
        return FITV2_OK;
    }
    const int q_pairs = ((tokens + 127) / 128 + 1) / 2;
    const int items = q_pairs * c.num_heads * rows;
    const int g = items < h->num_sms ? items : h->num_sms;
    int rc;
    if (mode == 4 && c.head_dim == 72) {                                        // experiment: one softmax thread per query row
        using A = AttnTmCfg<72, 1>;
        auto kern = attention_tm_kernel<OT, 72, 1>;
        const int smem = A::smem_bytes(tokens);
        if (smem > kSmemBudget) return fail(FITV2_E_INVALID, "tokens %d: attention shared memory %d exceeds %d", tokens, smem, kSmemBudget);
        if ((rc = ensure_smem(h, kern, smem))) return rc;
        CUDA_TRY(launch_k(kern, dim3(g), dim3(A::kThreads), smem, st, 1, am.mq, am.mqt, am.mk, am.mkt, am.mv, am.mo, seg, seg_uniform,
                          c.num_heads, tokens, items, scale_log2e, bound_log2e, h->opt.attn_early));
    } else if (mode == 1) {
        using A = AttnTmCfg<72>;
        auto kern = attention_tm_kernel<OT, 72>;
        const int smem = A::smem_bytes(tokens);
        if (smem > kSmemBudget) return fail(FITV2_E_INVALID, "tokens %d: attention shared memory %d exceeds %d", tokens, smem, kSmemBudget);
        if ((rc = ensure_smem(h, kern, smem))) return rc;
        // softmax of the next work item's first sub-tile ahead of the epilogue of the current one: 43.9 -> 40.8 us alone,
        // 57.7 -> 53.2 us in-step at 256 tokens (A/B on one box); option attn_early = 0 restores the plain order
        CUDA_TRY(launch_k(kern, dim3(g), dim3(A::kThreads), smem, st, 1, am.mq, am.mqt, am.mk, am.mkt, am.mv, am.mo, seg, seg_uniform,
                          c.num_heads, tokens, items, scale_log2e, bound_log2e, h->opt.attn_early));
    } else if (c.head_dim == 72) {
        using A = AttnWsCfg<72>;
        auto kern = attention_ws_kernel<OT, 72>;
        const int smem = A::smem_bytes(tokens);
        if (smem > kSmemBudget) return fail(FITV2_E_INVALID, "tokens %d: attention shared memory %d exceeds %d", tokens, smem, kSmemBudget);
        if ((rc = ensure_smem(h, kern, smem))) return rc;
        CUDA_TRY(launch_k(kern, dim3(g), dim3(A::kThreads), smem, st, 1, am.mq, am.mqt, am.mk, am.mkt, am.mv, am.mo, seg, seg_uniform,
                          c.num_heads, tokens, items, scale_log2e, bound_log2e));
    } else {
        using A = AttnWsCfg<96>;
        auto kern = attention_ws_kernel<OT, 96>;
        const int smem = A::smem_bytes(tokens);
        if (smem > kSmemBudget) return fail(FITV2_E_INVALID, "tokens %d: attention shared memory %d exceeds %d", tokens, smem, kSmemBudget);
        if ((rc = ensure_smem(h, kern, smem))) return rc;
        CUDA_TRY(launch_k(kern, dim3(g), dim3(A::kThreads), smem, st, 1, am.mq, am.mqt, am.mk, am.mkt, am.mv, am.mo, seg, seg_uniform,
                          c.num_heads, tokens, items, scale_log2e, bound_log2e));
    }
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return FITV2_OK;
}

// LayerNorm / RMSNorm (+ weight) + adaLN modulate -> 16-bit operand.  norm_mode: FITV2_NORM_* of the block norms.
template <typename OT>
int launch_ln_modulate(fitv2_handle* h, const float* x, const float* shift, const float* scale, int mod_ld, void* out,
                       int M, int D, int tokens, cudaStream_t st, int norm_mode = FITV2_NORM_LAYERNORM, const float* norm_w = nullptr) {
    const int nv = (D / 4 + 31) / 32;
    // two rows per CTA: measured 25.8 us per launch against 26.2 (4 rows) and 28.0 (8 rows) - shorter tail, finer CTA refill
    const int lt = h->opt.ln_threads;
    const int ln_threads = (lt == 32 || lt == 64 || lt == 128 || lt == 256) ? lt : 64;
    const int blocks = (M * 32 + ln_threads - 1) / ln_threads;
    if (nv > 18) return fail(FITV2_E_INVALID, "hidden_size %d too large for the LayerNorm kernel", D);
    if (norm_mode != FITV2_NORM_LAYERNORM) {
        if (norm_mode == FITV2_NORM_NONE) return fail(FITV2_E_INVALID, "block norm 'none' is not supported");
        if (!norm_w) return fail(FITV2_E_UNBOUND, "block norm weight is not bound");
#define FITV2_LN_W(NV_)                                                                                                                           \
        do {                                                                                                                                      \
            if (norm_mode == FITV2_NORM_WLAYERNORM)                                                                                               \
                CUDA_TRY(launch_k(ln_modulate_kernel<OT, NV_, 1>, dim3(blocks), dim3(ln_threads), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens, norm_w)); \
            else                                                                                                                                  \
                CUDA_TRY(launch_k(ln_modulate_kernel<OT, NV_, 2>, dim3(blocks), dim3(ln_threads), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens, norm_w)); \
        } while (0)
        if (nv <= 9) FITV2_LN_W(9); else FITV2_LN_W(18);
#undef FITV2_LN_W
    }
    else if (nv <= 1) CUDA_TRY(launch_k(ln_modulate_kernel<OT, 1>, dim3(blocks), dim3(ln_threads), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens, norm_w));
    else if (nv <= 3) CUDA_TRY(launch_k(ln_modulate_kernel<OT, 3>, dim3(blocks), dim3(ln_threads), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens, norm_w));
    else if (D == 1152) CUDA_TRY(launch_k(ln_modulate_kernel<OT, 9, 0, true>, dim3(blocks), dim3(ln_threads), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens, norm_w));
    else if (nv <= 9) CUDA_TRY(launch_k(ln_modulate_kernel<OT, 9>, dim3(blocks), dim3(ln_threads), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens, norm_w));
    else {                                                           // wide rows: one row per warp pair (4 rows per CTA)
        if (h->opt.ln_wide_single) CUDA_TRY(launch_k(ln_modulate_kernel<OT, 18>, dim3(blocks), dim3(ln_threads), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens, norm_w));
        else CUDA_TRY(launch_k(ln_modulate_pair_kernel<OT, 9>, dim3((M + 3) / 4), dim3(256), 0, st, 1, x, shift, scale, mod_ld, (OT*)out, M, D, tokens));
    }
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return FITV2_OK;
}

int launch_small_linear(fitv2_handle* h, SmallLinear p, int batches, cudaStream_t st) {
    const int col_tiles = (p.N + 63) / 64, row_blocks = (p.rows + 63) / 64;
    const long ctas = (long)col_tiles * row_blocks * batches;
    // fill the GPU: split K when the output alone gives fewer than ~2 CTAs per SM (each split keeps >= 64 of K)
    long ksplit = (2L * h->num_sms + ctas - 1) / ctas;
    const long max_by_k = p.K / 64 > 0 ? p.K / 64 : 1;
    const long max_by_mem = (long)(kCondPartialBytes / ((size_t)batches * p.rows * p.N * 4));
    if (ksplit > max_by_k) ksplit = max_by_k;
    if (ksplit > max_by_mem) ksplit = max_by_mem;
    if (ksplit < 1) ksplit = 1;
    p.ksplit = (int)ksplit;
    p.partial = reinterpret_cast<float*>(h->ws + h->lay.cpart);
    dim3 grid(col_tiles, row_blocks * p.ksplit, batches);
    CUDA_TRY(launch_k(small_linear_kernel, grid, dim3(256), 0, st, 1, p));
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    if (p.ksplit > 1) {
        dim3 g2((unsigned)(((size_t)p.rows * p.N + 255) / 256), 1, batches);
        CUDA_TRY(launch_k(small_linear_finalize_kernel, g2, dim3(256), 0, st, 1, p));
        CUDA_TRY(cudaGetLastError());
        h->launches++;
    }
    return FITV2_OK;
}

int ensure_maps(fitv2_handle* h) {
    if (h->maps_valid) return FITV2_OK;
    const fitv2_config& c = h->cfg;
    const Options& o = h->opt;
    const Layout& l = h->lay;
    const uint64_t M = (uint64_t)l.rows * l.tokens, D = c.hidden_size, Hm = c.mlp_hidden, L = c.depth;
    int rc;
    h->attn_maps = AttnMaps();
    if ((rc = make_map(&h->map_h, h->ws + l.h, c.operand_dtype, M, D, D, 128))) return rc;
    if ((rc = make_map(&h->map_ao, h->ws + l.ao, c.operand_dtype, M, D, D, 128))) return rc;
    if ((rc = make_map(&h->map_hidden, h->ws + l.hidden, c.operand_dtype, M, Hm, Hm, 128))) return rc;
    // proj (K = D) is bound by its fp32 residual epilogue: prefer tile widths whose halves start on 128-byte lines;
    // fc2 (K = mlp_hidden) has the longer main loop: pure wave / tile-size cost.
    h->bn_proj = pick_bn((int)M, (int)D, h->num_sms, /*prefer_aligned=*/true);
    h->bn_fc2 = pick_bn((int)M, (int)D, h->num_sms, /*prefer_aligned=*/false);
    if ((o.bn_resid == 128 || o.bn_resid == 144 || o.bn_resid == 192 || o.bn_resid == 256) && D % o.bn_resid == 0)
        h->bn_proj = h->bn_fc2 = o.bn_resid;                        // tuning experiments only
    if (!h->bn_proj || !h->bn_fc2) return fail(FITV2_E_INVALID, "hidden_size %d has no supported tile width (multiple of 128/144/192/256)", (int)D);
    // q / k norm other than the affine-free LayerNorm of the FiTv2 configs: generic epilogue (two-head tile) + online-max attention
    // (rope_v, the rotation of v, also lives in the generic epilogue; the attention kernel choice only follows the q / k norms)
    h->qkv_gen = !(c.q_norm == FITV2_NORM_LAYERNORM && c.k_norm == FITV2_NORM_LAYERNORM) || c.rope_v != 0;
    // head_dim 72: three heads per 224-wide tile (option qkv_heads = 2 keeps the two-head 144-wide tile)
    h->qkv3 = c.head_dim == 72 && o.qkv_heads != 2 && !h->qkv_gen;
    if ((rc = make_map(&h->map_wqkv, h->w[FITV2_W_QKV_W], c.operand_dtype, L * 3 * D, D, D,
                       (h->qkv3 ? 3 * c.head_dim + 8 : 2 * c.head_dim) / kGemmCluster))) return rc;
    if ((rc = make_map(&h->map_wproj, h->w[FITV2_W_PROJ_W], c.operand_dtype, L * D, D, D, h->bn_proj / kGemmCluster))) return rc;
    if ((rc = make_map(&h->map_wgu, h->w[FITV2_W_GATEUP_W], c.operand_dtype, L * (c.mlp_type == FITV2_MLP_GELU ? 1 : 2) * Hm, D, D, 256 / kGemmCluster))) return rc;
    if ((rc = make_map(&h->map_wfc2, h->w[FITV2_W_FC2_W], c.operand_dtype, L * D, Hm, Hm, h->bn_fc2 / kGemmCluster))) return rc;
    {
        // Transposed residual GEMM (EPI_RESID_T): 256-token-wide tiles + TMA reduce-add into x.  Measured at XL/2 (hidden 1152,
        // five 256-channel groups, the last half empty): fc2 116 -> 110 us, proj 61 -> 63 us (proj is bound by the DRAM traffic of
        // the residual either way), so only fc2 uses it, and only where the normal orientation has no 256-wide tile.
        // Option resid_t = 0 / 1 forces it off / on for both.
        // Few token rows (batch 2: 1 024 rows = 20 transposed tiles for 74 clusters, 27 us per launch): the narrower normal-orientation
        // tiles fill more clusters, decided by the same per-K-block cost model below.
        h->proj_t = o.resid_t == 1;
        if ((rc = make_map(&h->map_wproj_t, h->w[FITV2_W_PROJ_W], c.operand_dtype, L * D, D, D, 128))) return rc;
        {
            EncodeTiledFn fn = get_encode_fn();
            if (!fn) return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
            cuuint64_t dims[2] = {D, M};
            cuuint64_t strides[1] = {D * 4};
            cuuint32_t box[2] = {32, 16};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = fn(&h->map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->ws + l.x_res, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(FITV2_E_CUDA, "cuTensorMapEncodeTiled(residual) failed (%d)", (int)r);
        }
        if ((rc = make_map(&h->map_wfc2_t, h->w[FITV2_W_FC2_W], c.operand_dtype, L * D, Hm, Hm, 128))) return rc;
        {   // token-tile width: waves of (channel groups x token tiles) over the clusters times the cost of one K block
            // (tensor pipe 2 * bn clocks, shared-memory pipe 256 + bn); 16 384 rows: 5 waves either way, 224 is 6 % cheaper
            const long groups = ((long)((D + kGemmBM - 1) / kGemmBM) + kGemmCluster - 1) / kGemmCluster, clusters = h->num_sms / kGemmCluster;
            long best = -1;
            const int cand[2] = {256, 224};
            for (int bn : cand) {
                const long tiles = groups * (((long)M + bn - 1) / bn);
                const long cost = ((tiles + clusters - 1) / clusters) * (2L * bn > 256L + bn ? 2L * bn : 256L + bn);
                if (best < 0 || cost < best) { best = cost; h->bn_resid_t = bn; }
            }
            if (o.bn_resid_t == 256 || o.bn_resid_t == 224) h->bn_resid_t = o.bn_resid_t;
            auto kblock_clocks = [](long bn) { return 2L * bn > 256L + bn ? 2L * bn : 256L + bn; };
            const long bt = h->bn_resid_t, bn = h->bn_fc2;
            const long cost_t = ((groups * (((long)M + bt - 1) / bt) + clusters - 1) / clusters) * kblock_clocks(bt);
            const long tiles_n = (((long)M + 2 * kGemmBM - 1) / (2 * kGemmBM)) * ((long)D / bn);
            const long cost_n = ((tiles_n + clusters - 1) / clusters) * kblock_clocks(bn);
            h->fc2_t = o.resid_t < 0 ? (h->bn_fc2 < 256 && cost_t <= cost_n) : o.resid_t >= 1;     // 2: fc2 only, whatever the model says
        }
        if ((rc = make_map(&h->map_ao_t, h->ws + l.ao, c.operand_dtype, M, D, D, h->bn_resid_t / kGemmCluster))) return rc;
        if ((rc = make_map(&h->map_hidden_t, h->ws + l.hidden, c.operand_dtype, M, Hm, Hm, h->bn_resid_t / kGemmCluster))) return rc;
    }
    // Conditioning linears on the tensor pipe when the shapes tile (every production config does); option cond = 1 keeps the
    // fp32-FMA kernels for A/B runs.  Other shapes use the SIMT kernels.
    {
        const bool lora_mode = c.adaln_type == FITV2_ADALN_LORA;
        const uint64_t lora = lora_mode ? c.lora_dim : 0, split_rows = (uint64_t)((l.rows + kCondRows - 1) / kCondRows) * 128;
        h->cond_tc = o.cond == 0 && c.adaln_type != FITV2_ADALN_SWIGLU && D % 32 == 0 && (6 * D) % 48 == 0 && (2 * D) % 48 == 0 &&
                     (!lora_mode || (lora % 32 == 0 && (L * lora) % 48 == 0));
        if (h->cond_tc) {
            h->cond_bn_up = (6 * D) % 128 == 0 ? 128 : 48;
            if ((rc = make_map_f32(&h->map_sc_split, h->ws + l.sc_split, split_rows, D, D, 128))) return rc;
            if ((rc = make_map_f32(&h->map_wfinal, h->w[FITV2_W_FINAL_ADALN_W], 2 * D, D, D, 48))) return rc;
            if (lora_mode) {
                if ((rc = make_map_f32(&h->map_lmid_split, h->ws + l.lmid_split, split_rows, L * lora, L * lora, 128))) return rc;
                if ((rc = make_map_f32(&h->map_wglobal, h->w[FITV2_W_GLOBAL_ADALN_W], 6 * D, D, D, 48))) return rc;
                if ((rc = make_map_f32(&h->map_wlora_a, h->w[FITV2_W_LORA_A_W], L * lora, D, D, 48))) return rc;
                if ((rc = make_map_f32(&h->map_wlora_b, h->w[FITV2_W_LORA_B_W], L * 6 * D, lora, lora, h->cond_bn_up))) return rc;
            } else {
                if ((rc = make_map_f32(&h->map_wnormal, h->w[FITV2_W_NORMAL_ADALN_W], L * 6 * D, D, D, h->cond_bn_up))) return rc;
            }
        }
    }
    // Final layer on the tensor pipe: LayerNorm + modulate kernel -> fp16 operand, skinny tcgen05 GEMM (N = 16 / 32) with the
    // fp16 copy of final_layer.linear.weight; fp16 (10-bit mantissa) rather than the handle's operand type because this
    // product IS the output.  Option final_tc = 0 keeps the fused fp32 SIMT kernel.
    h->final_tc = o.final_tc != 0;
    h->wfin16_valid = false;
    if (h->final_tc) {
        const uint64_t Co = out_channels(c);
        if ((rc = make_map(&h->map_hfinal, h->ws + l.h, FITV2_OPERAND_FP16, M, D, D, 128))) return rc;
        if ((rc = make_map(&h->map_wfinal_lin, h->ws + l.wfin16, FITV2_OPERAND_FP16, Co, D, D, (uint32_t)(Co / kGemmCluster)))) return rc;
    }
    h->maps_valid = true;
    return FITV2_OK;
}

template <int BN>
int launch_cond_tc(fitv2_handle* h, const CUtensorMap& ma, const CUtensorMap& w0, const CUtensorMap& w1, const CUtensorMap& w2,
                   const CondTc& p, cudaStream_t st) {
    auto kern = cond_tc_kernel<BN>;
    int rc = ensure_smem(h, kern, CondCfg<BN>::kSmemBytes);
    if (rc) return rc;
    const int m_tiles = (p.rows + kCondRows - 1) / kCondRows;
    const int tiles = p.batches * m_tiles * p.nt_prefix[p.nseg];
    const int grid = tiles < h->num_sms ? tiles : h->num_sms;
    CUDA_TRY(launch_k(kern, dim3(grid), dim3(192), CondCfg<BN>::kSmemBytes, st, 1, ma, w0, w1, w2, p));
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return FITV2_OK;
}

template <typename OT>
int forward_impl(fitv2_handle* h, const float* x_in, int x_rows, const float* t, const int64_t* y, const int64_t* grid,
                 const float* mask, float* out_user, int rows, int tokens, cudaStream_t st) {
    const fitv2_config& c = h->cfg;
    const Layout& l = h->lay;
    const int D = c.hidden_size, L = c.depth, Hm = c.mlp_hidden, H = c.num_heads, DH = c.head_dim;
    const bool lora_mode = c.adaln_type == FITV2_ADALN_LORA;
    const int lora = lora_mode ? c.lora_dim : 0;
    const int Cin = c.token_channels, Cout = out_channels(c);
    const int M = rows * tokens;
    uint8_t* ws = h->ws;
    float* x_res = (float*)(ws + l.x_res);
    float* te = (float*)(ws + l.te);
    float* t0 = (float*)(ws + l.t0);
    float* cc = (float*)(ws + l.c);
    float* sc = (float*)(ws + l.sc);
    float* lmid = (float*)(ws + l.lmid);
    float* gmod = (float*)(ws + l.gmod);
    float* mod = (float*)(ws + l.mod);
    float* fmod = (float*)(ws + l.fmod);
    float* rcos = (float*)(ws + l.rope_cos);
    float* rsin = (float*)(ws + l.rope_sin);
    int* segu = (int*)(ws + l.seg_uniform);
    int rc;
    g_pdl = h->opt.pdl != 0;
    struct WindowGuard { ~WindowGuard() { g_l2_window = L2Window(); g_pdl = true; } } window_guard;
    // the window is attached only to the kernels that read or write the residual (LayerNorm, proj, fc2); everything else
    // launches without an access policy
    L2Window x_window;
    if (h->l2_persist_bytes > 0) {
        const size_t xb = (size_t)M * D * 4;
        x_window.base = x_res;
        x_window.bytes = xb < (size_t)h->l2_window_max ? xb : (size_t)h->l2_window_max;
        const float r = (float)h->l2_persist_bytes / (float)x_window.bytes;
        x_window.hit_ratio = r < 1.f ? r : 1.f;
    }
    auto x_kernels = [&](bool on) { g_l2_window = on ? x_window : L2Window(); };

    // ---- use_sit = False callers hand in (rows, C, tokens): token-major copy first (fit_model.py:204) ----
    const float* x = x_in;
    float* out = out_user;
    if (c.channels_first) {
        float* xt = (float*)(ws + l.xt);
        const size_t n = (size_t)x_rows * tokens * Cin;
        CUDA_TRY(launch_k(transpose_inner_kernel, dim3((unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184)), dim3(256), 0, st, 1, x_in, xt, x_rows, Cin, tokens));
        h->launches++;
        x = xt;
        out = (float*)(ws + l.ot);
    }

    // ---- per-call tables: segment-uniformity flags, RoPE cos/sin (rope.py:308-333) ----
    prof_begin(h, PC_COND, st);
    CUDA_TRY(launch_k(seg_uniform_kernel, dim3(rows), dim3(128), 0, st, 1, mask, segu, tokens));
    {
        const size_t total = (size_t)M * (DH / 2);
        const int blocks = (int)((total + 255) / 256);
        if (h->online_fh && h->online_rows != rows)
            return fail(FITV2_E_INVALID, "online RoPE frequencies were set for %d rows, forward has %d", h->online_rows, rows);
        if (h->online_fh)
            CUDA_TRY(launch_k(rope_table_kernel, dim3(blocks), dim3(256), 0, st, 1, (const long long*)grid, h->online_fh, h->online_fw, DH / 4,
                              1.0f, rcos, rsin, rows, tokens, DH / 2));
        else
            CUDA_TRY(launch_k(rope_table_kernel, dim3(blocks), dim3(256), 0, st, 1, (const long long*)grid, (const float*)h->w[FITV2_W_ROPE_FREQS_H],
                              (const float*)h->w[FITV2_W_ROPE_FREQS_W], 0, c.rope_magnitude, rcos, rsin, rows, tokens, DH / 2));
    }
    // ---- conditioning (fit_model.py:202-209,218-219; modules.py:52-76,101-106,254-264,287-293) ----
    CUDA_TRY(launch_k(timestep_features_kernel, dim3((rows * 128 + 255) / 256), dim3(256), 0, st, 1, t, c.time_shifting, te, rows));
    CUDA_TRY(cudaGetLastError());
    h->launches += 3;
    SmallLinear p;
    memset(&p, 0, sizeof(p));
    p.rows = rows;
    // t0 = W0 te + b0
    p.A = te; p.lda = 256; p.W = (const float*)h->w[FITV2_W_T_MLP0_W]; p.bias = (const float*)h->w[FITV2_W_T_MLP0_B];
    p.out = t0; p.ldo = D; p.N = D; p.K = 256;
    if ((rc = launch_small_linear(h, p, 1, st))) return rc;
    // c = W2 silu(t0) + b2 + E[y] ; sc = silu(c)
    p.A = t0; p.lda = D; p.act_silu_in = 1; p.W = (const float*)h->w[FITV2_W_T_MLP2_W]; p.bias = (const float*)h->w[FITV2_W_T_MLP2_B];
    p.emb = (const float*)h->w[FITV2_W_Y_TABLE]; p.labels = (const long long*)y; p.num_emb = c.num_embeddings; p.err = h->err_dev;
    p.out = cc; p.out_silu = sc; p.N = D; p.K = D;
    p.out_silu_split = h->cond_tc ? (float*)(ws + l.sc_split) : nullptr;   // hi / lo tf32 split of silu(c) for cond_tc.cuh
    if ((rc = launch_small_linear(h, p, 1, st))) return rc;
    p.act_silu_in = 0; p.emb = nullptr; p.labels = nullptr; p.out_silu = nullptr; p.out_silu_split = nullptr; p.err = nullptr;
    if (h->cond_tc && lora_mode) {
        // tensor-pipe path (cond_tc.cuh): one launch for the three linears that read silu(c), one for the batched LoRA up
        float* lmid_split = (float*)(ws + l.lmid_split);
        CondTc q;
        memset(&q, 0, sizeof(q));
        q.nseg = 3; q.rows = rows; q.K = D; q.batches = 1;
        q.nt_prefix[0] = 0; q.nt_prefix[1] = 6 * D / 48; q.nt_prefix[2] = q.nt_prefix[1] + 2 * D / 48; q.nt_prefix[3] = q.nt_prefix[2] + L * lora / 48;
        q.bias[0] = (const float*)h->w[FITV2_W_GLOBAL_ADALN_B]; q.out[0] = gmod; q.ldo[0] = 6 * D;
        q.bias[1] = (const float*)h->w[FITV2_W_FINAL_ADALN_B]; q.out[1] = fmod; q.ldo[1] = 2 * D;
        q.bias[2] = (const float*)h->w[FITV2_W_LORA_A_B]; q.out[2] = lmid; q.ldo[2] = L * lora;
        q.out_split[2] = lmid_split; q.ld_split[2] = L * lora;
        if ((rc = launch_cond_tc<48>(h, h->map_sc_split, h->map_wglobal, h->map_wfinal, h->map_wlora_a, q, st))) return rc;
        CondTc u;
        memset(&u, 0, sizeof(u));
        u.nseg = 1; u.rows = rows; u.K = lora; u.batches = L; u.a_batch_cols = lora;
        u.nt_prefix[0] = 0; u.nt_prefix[1] = 6 * D / h->cond_bn_up;
        u.w_batch_rows[0] = 6 * D;
        u.bias[0] = (const float*)h->w[FITV2_W_LORA_B_B]; u.bias_batch_stride[0] = 6 * D;
        u.add[0] = gmod; u.out[0] = mod; u.out_batch_stride[0] = (size_t)rows * 6 * D; u.ldo[0] = 6 * D;
        if (h->cond_bn_up == 128) rc = launch_cond_tc<128>(h, h->map_lmid_split, h->map_wlora_b, h->map_wlora_b, h->map_wlora_b, u, st);
        else rc = launch_cond_tc<48>(h, h->map_lmid_split, h->map_wlora_b, h->map_wlora_b, h->map_wlora_b, u, st);
        if (rc) return rc;
    } else if (h->cond_tc) {
        // adaln_type 'normal' (modules.py:254-258): final adaLN, then every block's Linear(D -> 6D) on silu(c) as one batched launch
        CondTc q;
        memset(&q, 0, sizeof(q));
        q.nseg = 1; q.rows = rows; q.K = D; q.batches = 1;
        q.nt_prefix[0] = 0; q.nt_prefix[1] = 2 * D / 48;
        q.bias[0] = (const float*)h->w[FITV2_W_FINAL_ADALN_B]; q.out[0] = fmod; q.ldo[0] = 2 * D;
        if ((rc = launch_cond_tc<48>(h, h->map_sc_split, h->map_wfinal, h->map_wfinal, h->map_wfinal, q, st))) return rc;
        CondTc u;
        memset(&u, 0, sizeof(u));
        u.nseg = 1; u.rows = rows; u.K = D; u.batches = L; u.a_batch_cols = 0;
        u.nt_prefix[0] = 0; u.nt_prefix[1] = 6 * D / h->cond_bn_up;
        u.w_batch_rows[0] = 6 * D;
        u.bias[0] = (const float*)h->w[FITV2_W_NORMAL_ADALN_B]; u.bias_batch_stride[0] = 6 * D;
        u.out[0] = mod; u.out_batch_stride[0] = (size_t)rows * 6 * D; u.ldo[0] = 6 * D;
        if (h->cond_bn_up == 128) rc = launch_cond_tc<128>(h, h->map_sc_split, h->map_wnormal, h->map_wnormal, h->map_wnormal, u, st);
        else rc = launch_cond_tc<48>(h, h->map_sc_split, h->map_wnormal, h->map_wnormal, h->map_wnormal, u, st);
        if (rc) return rc;
    } else if (c.adaln_type == FITV2_ADALN_SWIGLU) {
        // modules.py:265-268,284-285: the modulation is a SwiGLU MLP on c ITSELF (no SiLU in front, no global term):
        // g = Wg c + bg (+ silu(g)), m = (Wx c + bx) * silu(g), out = W2 m + b2; batched over the blocks, then once for the final layer
        float* gb = (float*)(ws + l.sg_g);
        float* sb = (float*)(ws + l.sg_s);
        float* mb = (float*)(ws + l.sg_m);
        auto sg_mlp = [&](int batches, int Hh, int Nout, int slot0, float* outp) -> int {
            SmallLinear q;
            memset(&q, 0, sizeof(q));
            q.rows = rows; q.A = cc; q.lda = D; q.K = D; q.N = Hh; q.ldo = Hh; q.out_batch_stride = (size_t)rows * Hh;
            q.W = (const float*)h->w[slot0]; q.w_batch_stride = (size_t)Hh * D; q.bias = (const float*)h->w[slot0 + 1]; q.bias_batch_stride = Hh;
            q.out = gb; q.out_silu = sb;
            int r2 = launch_small_linear(h, q, batches, st);
            if (r2) return r2;
            q.W = (const float*)h->w[slot0 + 2]; q.bias = (const float*)h->w[slot0 + 3];
            q.out = mb; q.out_silu = nullptr; q.mul = sb; q.mul_batch_stride = (size_t)rows * Hh; q.ldm = Hh;
            if ((r2 = launch_small_linear(h, q, batches, st))) return r2;
            memset(&q, 0, sizeof(q));
            q.rows = rows; q.A = mb; q.a_batch_stride = (size_t)rows * Hh; q.lda = Hh; q.K = Hh; q.N = Nout; q.ldo = Nout;
            q.W = (const float*)h->w[slot0 + 4]; q.w_batch_stride = (size_t)Nout * Hh; q.bias = (const float*)h->w[slot0 + 5]; q.bias_batch_stride = Nout;
            q.out = outp; q.out_batch_stride = (size_t)rows * Nout;
            return launch_small_linear(h, q, batches, st);
        };
        if ((rc = sg_mlp(L, (D / 4) * 3, 6 * D, FITV2_W_SG_G_W, mod))) return rc;
        if ((rc = sg_mlp(1, D / 2, 2 * D, FITV2_W_FSG_G_W, fmod))) return rc;
    } else {
        // final adaLN: fmod = Wf sc + bf   (shift | scale)
        p.A = sc; p.lda = D; p.K = D;
        p.W = (const float*)h->w[FITV2_W_FINAL_ADALN_W]; p.bias = (const float*)h->w[FITV2_W_FINAL_ADALN_B];
        p.out = fmod; p.ldo = 2 * D; p.N = 2 * D;
        if ((rc = launch_small_linear(h, p, 1, st))) return rc;
        if (lora_mode) {
            // global adaLN: gmod = Wg sc + bg
            p.W = (const float*)h->w[FITV2_W_GLOBAL_ADALN_W]; p.bias = (const float*)h->w[FITV2_W_GLOBAL_ADALN_B];
            p.out = gmod; p.ldo = 6 * D; p.N = 6 * D;
            if ((rc = launch_small_linear(h, p, 1, st))) return rc;
            // LoRA down for all blocks at once: lmid = Wa_all sc + ba_all
            p.W = (const float*)h->w[FITV2_W_LORA_A_W]; p.bias = (const float*)h->w[FITV2_W_LORA_A_B];
            p.out = lmid; p.ldo = L * lora; p.N = L * lora;
            if ((rc = launch_small_linear(h, p, 1, st))) return rc;
            // LoRA up, batched over blocks: mod[l] = Wb[l] lmid[:, l] + bb[l] + gmod
            p.A = lmid; p.a_batch_stride = lora; p.lda = L * lora;
            p.W = (const float*)h->w[FITV2_W_LORA_B_W]; p.w_batch_stride = (size_t)6 * D * lora;
            p.bias = (const float*)h->w[FITV2_W_LORA_B_B]; p.bias_batch_stride = 6 * D;
            p.add = gmod; p.out = mod; p.out_batch_stride = (size_t)rows * 6 * D; p.ldo = 6 * D; p.N = 6 * D; p.K = lora;
            if ((rc = launch_small_linear(h, p, L, st))) return rc;
        } else {
            // adaln_type 'normal': mod[l] = W[l] sc + b[l], batched over blocks
            p.W = (const float*)h->w[FITV2_W_NORMAL_ADALN_W]; p.w_batch_stride = (size_t)6 * D * D;
            p.bias = (const float*)h->w[FITV2_W_NORMAL_ADALN_B]; p.bias_batch_stride = 6 * D;
            p.out = mod; p.out_batch_stride = (size_t)rows * 6 * D; p.ldo = 6 * D; p.N = 6 * D;
            if ((rc = launch_small_linear(h, p, L, st))) return rc;
        }
    }
    prof_end(h, st);

    // ---- patch embedding (modules.py:34-37); implicit cat([z, z]) when x_rows == rows / 2 ----
    prof_begin(h, PC_MISC, st);
    const int pe_threads = D / 4 >= 576 ? 576 : ((D / 4 + 31) / 32) * 32;   // more features than threads: the kernel loops
    const int pe_groups = (M + kPatchRows - 1) / kPatchRows, pe_grid = pe_groups < 2 * h->num_sms ? pe_groups : 2 * h->num_sms;   // persistent: two blocks are resident per SM (96 registers)
    CUDA_TRY(launch_k(patch_embed_kernel<16>, dim3(pe_grid), dim3(pe_threads), 0, st, 1, x, (const float*)h->w[FITV2_W_X_EMBED_W], (const float*)h->w[FITV2_W_X_EMBED_B],
                                                       x_res, M, D, x_rows * tokens));
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    prof_end(h, st);

    const bool norm_w = norm_has_weight(c.block_norm);
    GemmEpi ep;
    for (int layer = 0; layer < L; ++layer) {
        const float* modl = mod + (size_t)layer * rows * 6 * D;
        // ---- attention branch (modules.py:272) ----
        prof_begin(h, PC_LNMOD, st);
        x_kernels(true);
        if ((rc = launch_ln_modulate<OT>(h, x_res, modl, modl + D, 6 * D, ws + l.h, M, D, tokens, st, c.block_norm,
                                         norm_w ? (const float*)h->w[FITV2_W_NORM1_W] + (size_t)layer * D : nullptr))) return rc;
        x_kernels(false);
        prof_end(h, st);
        memset(&ep, 0, sizeof(ep));
        ep.bias = (const float*)h->w[FITV2_W_QKV_B] + (size_t)layer * 3 * D;
        ep.tokens = tokens; ep.q = ws + l.q; ep.k = ws + l.k; ep.vt = ws + l.vt; ep.rope_cos = rcos; ep.rope_sin = rsin;
        ep.heads = H; ep.tokens_v = l.tokens_v;
        prof_begin(h, PC_QKV, st);
        if (h->qkv_gen) {
            ep.q_norm = c.q_norm; ep.k_norm = c.k_norm; ep.rope_v = c.rope_v;
            ep.q_norm_w = norm_has_weight(c.q_norm) ? (const float*)h->w[FITV2_W_Q_NORM_W] + (size_t)layer * DH : nullptr;
            ep.k_norm_w = norm_has_weight(c.k_norm) ? (const float*)h->w[FITV2_W_K_NORM_W] + (size_t)layer * DH : nullptr;
            if (DH == 72) rc = launch_gemm_t<144, EPI_QKV_GEN, OT, 72>(h, h->map_h, h->map_wqkv, M, 3 * D, D, layer * 3 * D, ep, st);
            else          rc = launch_gemm_t<192, EPI_QKV_GEN, OT, 96>(h, h->map_h, h->map_wqkv, M, 3 * D, D, layer * 3 * D, ep, st);
        }
        else if (DH == 72 && h->qkv3) rc = launch_gemm_t<224, EPI_QKV, OT, 72>(h, h->map_h, h->map_wqkv, M, 3 * D, D, layer * 3 * D, ep, st);
        else if (DH == 72) rc = launch_gemm_t<144, EPI_QKV, OT, 72>(h, h->map_h, h->map_wqkv, M, 3 * D, D, layer * 3 * D, ep, st);
        else          rc = launch_gemm_t<192, EPI_QKV, OT, 96>(h, h->map_h, h->map_wqkv, M, 3 * D, D, layer * 3 * D, ep, st);
        if (rc) return rc;
        prof_end(h, st);
        prof_begin(h, PC_ATTN, st);
        if ((rc = launch_attention<OT>(h, ws + l.q, ws + l.k, ws + l.vt, mask, segu, ws + l.ao, rows, tokens, l.tokens_v, st))) return rc;
        prof_end(h, st);
        memset(&ep, 0, sizeof(ep));
        ep.bias = (const float*)h->w[FITV2_W_PROJ_B] + (size_t)layer * D;
        ep.tokens = tokens; ep.x = x_res; ep.gate = modl + 2 * D; ep.gate_ld = 6 * D;
        prof_begin(h, PC_PROJ, st);
        x_kernels(true);
        if (h->proj_t && h->bn_resid_t == 224) rc = launch_gemm_t<224, EPI_RESID_T, OT, 0>(h, h->map_wproj_t, h->map_ao_t, D, M, D, layer * D, ep, st);
        else if (h->proj_t) rc = launch_gemm_t<256, EPI_RESID_T, OT, 0>(h, h->map_wproj_t, h->map_ao_t, D, M, D, layer * D, ep, st);
        else rc = launch_gemm_bn<EPI_RESID, OT>(h, h->bn_proj, h->map_ao, h->map_wproj, M, D, D, layer * D, ep, st);
        x_kernels(false);
        if (rc) return rc;
        prof_end(h, st);
        // ---- SwiGLU branch (modules.py:273) ----
        prof_begin(h, PC_LNMOD, st);
        x_kernels(true);
        if ((rc = launch_ln_modulate<OT>(h, x_res, modl + 3 * D, modl + 4 * D, 6 * D, ws + l.h, M, D, tokens, st, c.block_norm,
                                         norm_w ? (const float*)h->w[FITV2_W_NORM2_W] + (size_t)layer * D : nullptr))) return rc;
        x_kernels(false);
        prof_end(h, st);
        memset(&ep, 0, sizeof(ep));
        ep.tokens = tokens; ep.out16 = ws + l.hidden; ep.ld_out = Hm;
        prof_begin(h, PC_GATEUP, st);
        if (c.mlp_type == FITV2_MLP_GELU) {            // timm Mlp (modules.py:253): hidden = gelu_tanh(fc1(h)), plain 256-wide tiles
            ep.bias = (const float*)h->w[FITV2_W_GATEUP_B] + (size_t)layer * Hm;
            if (h->opt.gelu_epi == 1) {
                ep.act_gelu = 1;
                rc = launch_gemm_t<256, EPI_PLAIN, OT, 0>(h, h->map_h, h->map_wgu, M, Hm, D, layer * Hm, ep, st);
            } else {
                rc = launch_gemm_t<256, EPI_GELU, OT, 0>(h, h->map_h, h->map_wgu, M, Hm, D, layer * Hm, ep, st);
            }
            if (rc) return rc;
        } else {
            ep.bias = (const float*)h->w[FITV2_W_GATEUP_B] + (size_t)layer * 2 * Hm;
            if ((rc = launch_gemm_t<256, EPI_SWIGLU, OT, 0>(h, h->map_h, h->map_wgu, M, 2 * Hm, D, layer * 2 * Hm, ep, st))) return rc;
        }
        prof_end(h, st);
        memset(&ep, 0, sizeof(ep));
        ep.bias = (const float*)h->w[FITV2_W_FC2_B] + (size_t)layer * D;
        ep.tokens = tokens; ep.x = x_res; ep.gate = modl + 5 * D; ep.gate_ld = 6 * D;
        prof_begin(h, PC_FC2, st);
        x_kernels(true);
        if (h->fc2_t && h->bn_resid_t == 224) rc = launch_gemm_t<224, EPI_RESID_T, OT, 0>(h, h->map_wfc2_t, h->map_hidden_t, D, M, Hm, layer * D, ep, st);
        else if (h->fc2_t) rc = launch_gemm_t<256, EPI_RESID_T, OT, 0>(h, h->map_wfc2_t, h->map_hidden_t, D, M, Hm, layer * D, ep, st);
        else rc = launch_gemm_bn<EPI_RESID, OT>(h, h->bn_fc2, h->map_hidden, h->map_wfc2, M, D, Hm, layer * D, ep, st);
        x_kernels(false);
        if (rc) return rc;
        prof_end(h, st);
    }

    // ---- final layer + output mask (modules.py:292-296, fit_model.py:230) ----
    prof_begin(h, PC_MISC, st);
    const float* nfw = norm_w ? (const float*)h->w[FITV2_W_NORM_FINAL_W] : nullptr;
    if (h->final_tc) {
        // LN + modulate -> fp16 operand (the block kernel), then out = (h W^T + b) * mask as a skinny tcgen05 GEMM
        if (!h->wfin16_valid) {                        // once per (bind, workspace, layout): the copy lives in the handle's workspace
            const size_t nw = (size_t)Cout * D;
            CUDA_TRY(launch_k(f32_to_f16_kernel, dim3((unsigned)((nw + 255) / 256)), dim3(256), 0, st, 1, (const float*)h->w[FITV2_W_FINAL_LINEAR_W],
                              (__half*)(ws + l.wfin16), nw));
            h->launches++;
            h->wfin16_valid = true;
        }
        if ((rc = launch_ln_modulate<__half>(h, x_res, fmod, fmod + D, 2 * D, ws + l.h, M, D, tokens, st, c.block_norm, nfw))) return rc;
        memset(&ep, 0, sizeof(ep));
        ep.bias = (const float*)h->w[FITV2_W_FINAL_LINEAR_B];
        ep.tokens = tokens; ep.out32 = out; ep.ld_out = Cout; ep.row_scale = mask;
        if (Cout == 16) rc = launch_gemm_t<16, EPI_PLAIN, __half, 0>(h, h->map_hfinal, h->map_wfinal_lin, M, Cout, D, 0, ep, st);
        else            rc = launch_gemm_t<32, EPI_PLAIN, __half, 0>(h, h->map_hfinal, h->map_wfinal_lin, M, Cout, D, 0, ep, st);
        if (rc) return rc;
    } else {
        const int nv = (D / 4 + 31) / 32;
        const size_t smem = (size_t)16 * D * 4;
        const int blocks = h->num_sms * 2;             // measured: 85 us with two waves of blocks against 97 us with one block per SM
        const int rms = c.block_norm == FITV2_NORM_RMSNORM;
        for (int col0 = 0; col0 < Cout; col0 += 16) {  // learn_sigma: two launches of 16 output channels each
            if (nv <= 9) {
                auto kern = final_layer_kernel<9, 16>;
                if ((rc = ensure_smem(h, kern, 16 * 1152 * 4))) return rc;
                CUDA_TRY(launch_k(kern, dim3(blocks), dim3(256), smem, st, 1, x_res, fmod, (const float*)h->w[FITV2_W_FINAL_LINEAR_W], (const float*)h->w[FITV2_W_FINAL_LINEAR_B],
                                  mask, out, M, D, tokens, nfw, rms, Cout, col0));
            } else {
                auto kern = final_layer_kernel<18, 16>;
                if ((rc = ensure_smem(h, kern, 16 * 2304 * 4))) return rc;
                CUDA_TRY(launch_k(kern, dim3(blocks), dim3(256), smem, st, 1, x_res, fmod, (const float*)h->w[FITV2_W_FINAL_LINEAR_W], (const float*)h->w[FITV2_W_FINAL_LINEAR_B],
                                  mask, out, M, D, tokens, nfw, rms, Cout, col0));
            }
            CUDA_TRY(cudaGetLastError());
            h->launches++;
        }
    }
    if (c.channels_first) {                            // (rows, tokens, C_out) -> (rows, C_out, tokens)  (fit_model.py:231)
        const size_t n = (size_t)M * Cout;
        CUDA_TRY(launch_k(transpose_inner_kernel, dim3((unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184)), dim3(256), 0, st, 1, (const float*)out, out_user, rows, tokens, Cout));
        h->launches++;
    }
    prof_end(h, st);
    return FITV2_OK;
}

}  // namespace

// =================================================================================================
// exported C ABI
// =================================================================================================
extern "C" {

const char* fitv2_last_error(void) { return g_last_error.c_str(); }
const char* fitv2_version(void) { return "fitv2_b200 0.1 (sm_100a, tcgen05/TMEM/TMA)"; }

int fitv2_create(const fitv2_config* cfg, fitv2_handle** out) {
    if (!cfg || !out) return fail(FITV2_E_INVALID, "null argument");
    const fitv2_config& c = *cfg;
    if (c.hidden_size <= 0 || c.depth <= 0 || c.num_heads <= 0 || c.hidden_size != c.num_heads * c.head_dim)
        return fail(FITV2_E_INVALID, "hidden_size %d != num_heads %d * head_dim %d", c.hidden_size, c.num_heads, c.head_dim);
    if (c.head_dim != 72 && c.head_dim != 96)
        return fail(FITV2_E_INVALID, "head_dim %d not supported (kernels are built for 72 and 96)", c.head_dim);
    if (c.num_heads % 2) return fail(FITV2_E_INVALID, "num_heads %d must be even (two heads per QKV tile)", c.num_heads);
    if (c.mlp_hidden % 128) return fail(FITV2_E_INVALID, "mlp_hidden %d must be a multiple of 128", c.mlp_hidden);
    if (c.mlp_type != FITV2_MLP_SWIGLU && c.mlp_type != FITV2_MLP_GELU) return fail(FITV2_E_INVALID, "mlp_type %d unknown", c.mlp_type);
    if (c.mlp_type == FITV2_MLP_GELU && c.mlp_hidden % 256) return fail(FITV2_E_INVALID, "mlp_hidden %d must be a multiple of 256 for the GELU Mlp", c.mlp_hidden);
    if (c.rope_v != 0 && c.rope_v != 1) return fail(FITV2_E_INVALID, "rope_v %d must be 0 or 1", c.rope_v);
    if (c.hidden_size > 2304 || c.hidden_size % 16) return fail(FITV2_E_INVALID, "hidden_size %d must be a multiple of 16 and <= 2304", c.hidden_size);
    if (c.token_channels != 16) return fail(FITV2_E_INVALID, "token_channels %d not supported (p*p*C_in = 16)", c.token_channels);
    if (c.out_channels != 0 && c.out_channels != 16 && c.out_channels != 32)
        return fail(FITV2_E_INVALID, "out_channels %d not supported (16, or 32 with learn_sigma)", c.out_channels);
    if (c.adaln_type != FITV2_ADALN_LORA && c.adaln_type != FITV2_ADALN_NORMAL && c.adaln_type != FITV2_ADALN_SWIGLU)
        return fail(FITV2_E_INVALID, "adaln_type %d unknown", c.adaln_type);
    if (c.adaln_type == FITV2_ADALN_LORA && (c.lora_dim <= 0 || c.lora_dim % 4)) return fail(FITV2_E_INVALID, "lora_dim %d must be a positive multiple of 4", c.lora_dim);
    if (c.block_norm < FITV2_NORM_LAYERNORM || c.block_norm > FITV2_NORM_RMSNORM) return fail(FITV2_E_INVALID, "block_norm %d unknown (layernorm / w_layernorm / rmsnorm)", c.block_norm);
    if (c.q_norm < FITV2_NORM_NONE || c.q_norm > FITV2_NORM_RMSNORM || c.k_norm < FITV2_NORM_NONE || c.k_norm > FITV2_NORM_RMSNORM)
        return fail(FITV2_E_INVALID, "q_norm %d / k_norm %d unknown", c.q_norm, c.k_norm);
    if (c.operand_dtype != FITV2_OPERAND_BF16 && c.operand_dtype != FITV2_OPERAND_FP16)
        return fail(FITV2_E_INVALID, "operand_dtype %d unknown", c.operand_dtype);
    int dev = 0, major = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) return fail(FITV2_E_INVALID, "device compute capability %d.x: this library contains sm_100a code only", major);
    fitv2_handle* h = new fitv2_handle();
    h->cfg = c;
    h->device = dev;
    memset(h->w, 0, sizeof(h->w));
    memset(h->w_numel, 0, sizeof(h->w_numel));
    h->num_sms = sms;
    // sticky error word, written by kernels through the mapped device alias, read by the host without synchronising
    if (cudaHostAlloc(&h->err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&h->err_dev, h->err_host, 0) != cudaSuccess) {
        cudaGetLastError();
        if (h->err_host) cudaFreeHost(h->err_host);
        delete h;
        return fail(FITV2_E_CUDA, "cannot allocate the pinned error word");
    }
    *h->err_host = 0;
    *out = h;
    return FITV2_OK;
}

void fitv2_destroy(fitv2_handle* h) {
    if (!h) return;
    for (auto& e : h->prof_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (h->err_host) cudaFreeHost(h->err_host);
    delete h;
}

int fitv2_set_option(fitv2_handle* h, const char* name, int64_t value) {
    if (!h || !name) return fail(FITV2_E_INVALID, "null argument");
    Options& o = h->opt;
    const int v = (int)value;
    struct { const char* n; int* p; } tab[] = {
        {"pdl", &o.pdl}, {"attn", &o.attn}, {"attn_early", &o.attn_early}, {"ln_threads", &o.ln_threads},
        {"ln_wide_single", &o.ln_wide_single}, {"bn_resid", &o.bn_resid}, {"qkv_heads", &o.qkv_heads}, {"resid_t", &o.resid_t},
        {"bn_resid_t", &o.bn_resid_t}, {"cond", &o.cond}, {"l2_persist_mb", &o.l2_persist_mb}, {"final_tc", &o.final_tc}, {"gelu_epi", &o.gelu_epi}, {"ws_guard", &o.ws_guard}, {"verbose", &o.verbose}};
    for (auto& e : tab) {
        if (strcmp(e.n, name)) continue;
        *e.p = v;
        h->maps_valid = false;
        if (e.p == &o.ws_guard) h->lay = Layout();                    // buffer offsets change: re-derive at the next forward                                         // kernel selection / tile widths are derived in ensure_maps
        if (e.p == &o.l2_persist_mb) {                                 // persisting L2 for the fp32 residual stream (default off)
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
            h->l2_persist_bytes = 0;
            if (v > 0 && max_persist > 0) {
                long bytes = (long)v << 20;
                if (bytes > max_persist) bytes = max_persist;
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)bytes) == cudaSuccess) {
                    h->l2_persist_bytes = bytes;
                    h->l2_window_max = max_window;
                }
            }
            if (o.verbose) fprintf(stderr, "[fitv2] L2 persist max %d MB, window max %d MB, using %lld MB\n", max_persist >> 20, max_window >> 20, (long long)(h->l2_persist_bytes >> 20));
        }
        return FITV2_OK;
    }
    return fail(FITV2_E_INVALID, "unknown option '%s'", name);
}

int fitv2_poll_error(fitv2_handle* h) {
    if (!h) return fail(FITV2_E_INVALID, "null handle");
    const int e = *(volatile int*)h->err_host;
    if (e) {
        *(volatile int*)h->err_host = 0;
        return fail(FITV2_E_INVALID, "device-side check failed (flags 0x%x): a class label was outside [0, %d) (the reference raises an "
                                     "index error; row 0 of the table was used instead)", e, h->cfg.num_embeddings);
    }
    return FITV2_OK;
}

int fitv2_bind_weight(fitv2_handle* h, int slot, const void* dev_ptr, int64_t numel) {
    if (!h || slot < 0 || slot >= FITV2_W_COUNT || !dev_ptr) return fail(FITV2_E_INVALID, "bad bind_weight argument (slot %d)", slot);
    const int64_t want = expected_numel(h->cfg, slot);
    if (want == 0) return fail(FITV2_E_INVALID, "weight slot %d is not used by this configuration", slot);
    if (numel != want) return fail(FITV2_E_INVALID, "weight slot %d: got %lld elements, expected %lld", slot, (long long)numel, (long long)want);
    if (reinterpret_cast<uintptr_t>(dev_ptr) % 16) return fail(FITV2_E_INVALID, "weight slot %d pointer is not 16-byte aligned", slot);
    h->w[slot] = dev_ptr;
    h->w_numel[slot] = numel;
    h->maps_valid = false;
    return FITV2_OK;
}

int fitv2_set_online_rope(fitv2_handle* h, const float* freqs_h_rows, const float* freqs_w_rows, int rows) {
    if (!h) return fail(FITV2_E_INVALID, "null handle");
    if ((freqs_h_rows == nullptr) != (freqs_w_rows == nullptr) || (freqs_h_rows && rows <= 0))
        return fail(FITV2_E_INVALID, "bad online RoPE argument");
    h->online_fh = freqs_h_rows;
    h->online_fw = freqs_w_rows;
    h->online_rows = freqs_h_rows ? rows : 0;
    return FITV2_OK;
}

int64_t fitv2_workspace_bytes(const fitv2_handle* h, int rows, int tokens) {
    if (!h || rows <= 0 || tokens <= 0) return fail(FITV2_E_INVALID, "bad workspace query");
    return (int64_t)make_layout(h->cfg, rows, tokens, (size_t)h->opt.ws_guard).total;
}

int fitv2_set_workspace(fitv2_handle* h, void* dev_ptr, int64_t bytes) {
    if (!h || !dev_ptr || bytes <= 0) return fail(FITV2_E_INVALID, "bad workspace");
    if (reinterpret_cast<uintptr_t>(dev_ptr) % 256) return fail(FITV2_E_INVALID, "workspace must be 256-byte aligned");
    h->ws = static_cast<uint8_t*>(dev_ptr);
    h->ws_bytes = bytes;
    h->lay = Layout();
    h->maps_valid = false;
    return FITV2_OK;
}

int fitv2_forward(fitv2_handle* h, const float* x, int x_rows, const float* t, const int64_t* y, const int64_t* grid,
                  const float* mask, float* out, int rows, int tokens, void* stream) {
    if (!h || !x || !t || !y || !grid || !mask || !out) return fail(FITV2_E_INVALID, "null argument");
    if (rows <= 0 || tokens <= 0) return fail(FITV2_E_INVALID, "rows %d / tokens %d must be positive", rows, tokens);
    if (x_rows != rows && !(rows % 2 == 0 && x_rows == rows / 2))
        return fail(FITV2_E_INVALID, "x_rows %d must equal rows %d or rows/2", x_rows, rows);
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != h->device) return fail(FITV2_E_INVALID, "handle was created on device %d, the current device is %d", h->device, dev);
    for (int s = 0; s < FITV2_W_COUNT; ++s)
        if (!h->w[s] && expected_numel(h->cfg, s) > 0) return fail(FITV2_E_UNBOUND, "weight slot %d is not bound", s);
    if (!h->ws) return fail(FITV2_E_UNBOUND, "workspace is not set");
    int rc = fitv2_poll_error(h);                                      // an earlier launch saw an out-of-range label
    if (rc) return rc;
    if (h->lay.rows != rows || h->lay.tokens != tokens) {
        Layout l = make_layout(h->cfg, rows, tokens, (size_t)h->opt.ws_guard);
        if ((int64_t)l.total > h->ws_bytes)
            return fail(FITV2_E_WORKSPACE, "workspace has %lld bytes, (rows=%d, tokens=%d) needs %lld", (long long)h->ws_bytes, rows, tokens, (long long)l.total);
        h->lay = l;
        h->maps_valid = false;
    }
    rc = ensure_maps(h);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->cfg.operand_dtype == FITV2_OPERAND_FP16)
        return forward_impl<__half>(h, x, x_rows, t, y, grid, mask, out, rows, tokens, st);
    return forward_impl<__nv_bfloat16>(h, x, x_rows, t, y, grid, mask, out, rows, tokens, st);
}

int fitv2_cfg_combine(float* out, const float* scale_per_sample, float scale, int half_rows, int tokens, int channels,
                      int c_cfg, void* stream) {
    if (!out || half_rows <= 0 || tokens <= 0 || channels <= 0 || c_cfg < 0 || c_cfg > channels)
        return fail(FITV2_E_INVALID, "bad cfg_combine argument");
    const size_t total = (size_t)half_rows * tokens * channels;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    CUDA_TRY(launch_k(cfg_combine_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, out, scale_per_sample, scale, half_rows, tokens, channels, c_cfg));
    CUDA_TRY(cudaGetLastError());
    return FITV2_OK;
}

int fitv2_cfg_euler(float* z, const float* v2, float cfg_scale, float dsigma, const float* dsigma_dev, int half_rows,
                    int tokens, int channels, void* stream) {
    if (!z || !v2 || half_rows <= 0 || tokens <= 0 || channels <= 0) return fail(FITV2_E_INVALID, "bad cfg_euler argument");
    const size_t half = (size_t)half_rows * tokens * channels;
    if ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(v2) | (half * 4)) % 16)
        return fail(FITV2_E_INVALID, "cfg_euler needs 16-byte aligned z / v2 and a multiple of 4 elements per half");
    const size_t nvec = half / 4;
    const int blocks = (int)((nvec + 255) / 256 < 1184 ? (nvec + 255) / 256 : 1184);
    CUDA_TRY(launch_k(cfg_euler_kernel, dim3(blocks > 0 ? blocks : 1), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, z, v2, cfg_scale, dsigma, dsigma_dev, half));
    CUDA_TRY(cudaGetLastError());
    return FITV2_OK;
}

// ---- transport sampler updates (elementwise, fp32, bit-exact with the reference expressions) ----
namespace {
int elementwise_grid(int64_t n) {
    const int64_t blocks = (n / 4 + 255) / 256;
    return (int)(blocks < 1 ? 1 : (blocks > 1184 ? 1184 : blocks));
}
bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }
}  // namespace

int fitv2_sde_step(float* x, const float* v, const float* w, const float* coef_dev, int64_t n, void* stream) {
    if (!x || !v || !coef_dev || n <= 0) return fail(FITV2_E_INVALID, "bad sde_step argument");
    if (!aligned16(x) || !aligned16(v) || (w && !aligned16(w))) return fail(FITV2_E_INVALID, "sde_step needs 16-byte aligned tensors");
    CUDA_TRY(launch_k(sde_em_step_kernel, dim3(elementwise_grid(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, x, v, w, coef_dev, (size_t)n));
    return FITV2_OK;
}

int fitv2_sde_drift(float* out, const float* x, const float* v, const float* coef_dev, int64_t n, void* stream) {
    if (!out || !x || !v || !coef_dev || n <= 0) return fail(FITV2_E_INVALID, "bad sde_drift argument");
    CUDA_TRY(launch_k(sde_drift_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, out, x, v, coef_dev, (size_t)n));
    return FITV2_OK;
}

int fitv2_scaled_add(float* out, const float* a, const float* b, const float* s_dev, int64_t n, void* stream) {
    if (!out || !a || !b || !s_dev || n <= 0) return fail(FITV2_E_INVALID, "bad scaled_add argument");
    CUDA_TRY(launch_k(scaled_add_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, out, a, b, s_dev, (size_t)n));
    return FITV2_OK;
}

int fitv2_heun_combine(float* out, const float* xhat, const float* k1, const float* k2, const float* c_dev, int64_t n, void* stream) {
    if (!out || !xhat || !k1 || !k2 || !c_dev || n <= 0) return fail(FITV2_E_INVALID, "bad heun_combine argument");
    CUDA_TRY(launch_k(heun_combine_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, out, xhat, k1, k2, c_dev, (size_t)n));
    return FITV2_OK;
}

int fitv2_tweedie(float* out, const float* x, const float* v, const float* coef_dev, int64_t n, void* stream) {
    if (!out || !x || !v || !coef_dev || n <= 0) return fail(FITV2_E_INVALID, "bad tweedie argument");
    CUDA_TRY(launch_k(tweedie_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, out, x, v, coef_dev, (size_t)n));
    return FITV2_OK;
}

int fitv2_rk_stage(float* out, const float* y, const float* k1, const float* k2, const float* k3, const float* k4,
                   const float* s_dev, int mode, int64_t n, void* stream) {
    if (!out || !y || !k1 || !s_dev || n <= 0 || mode < 0 || mode > 5) return fail(FITV2_E_INVALID, "bad rk_stage argument");
    if ((mode >= 1 && !k2) || ((mode == 2 || mode >= 4) && !k3) || (mode == 5 && !k4)) return fail(FITV2_E_INVALID, "rk_stage mode %d misses a slope", mode);
    CUDA_TRY(launch_k(rk_stage_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, out, y, k1, k2, k3, k4, s_dev, mode, (size_t)n));
    return FITV2_OK;
}

int fitv2_lincomb(float* out, const float* y, const float* const* k, const float* c_dev, int nk, int64_t n, void* stream) {
    if (!out || !y || !c_dev || nk < 0 || nk > 7 || (nk > 0 && !k) || n <= 0) return fail(FITV2_E_INVALID, "bad lincomb argument");
    LinComb ks;
    for (int i = 0; i < 7; ++i) {
        ks.k[i] = i < nk ? k[i] : y;
        if (!ks.k[i]) return fail(FITV2_E_INVALID, "lincomb: slope %d is null", i);
    }
    CUDA_TRY(launch_k(lincomb_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, out, y, ks, c_dev, nk, (size_t)n));
    return FITV2_OK;
}

int fitv2_scaled_rms(float* out_dev, const float* a, const float* b, const float* s, float atol, float rtol, int64_t n, void* stream) {
    if (!out_dev || !a || n <= 0) return fail(FITV2_E_INVALID, "bad scaled_rms argument");
    CUDA_TRY(launch_k(scaled_rms_kernel, dim3(kRmsCluster), dim3(1024), 0, static_cast<cudaStream_t>(stream), kRmsCluster, out_dev, a, b, s, atol, rtol, (size_t)n));
    return FITV2_OK;
}

int fitv2_unpatchify_scale(const float* z, float* out, float scaling_factor, int batch, int hp, int wp, int channels, int patch, void* stream) {
    if (!z || !out || batch <= 0 || hp <= 0 || wp <= 0 || channels <= 0 || patch <= 0 || scaling_factor == 0.0f)
        return fail(FITV2_E_INVALID, "bad unpatchify_scale argument");
    const int64_t n = (int64_t)batch * channels * hp * wp * patch * patch;
    CUDA_TRY(launch_k(unpatchify_scale_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                      z, out, scaling_factor, batch, hp, wp, channels, patch));
    return FITV2_OK;
}

int fitv2_pack_uint8(const float* img, unsigned char* out, int batch, int channels, int height, int width, void* stream) {
    if (!img || !out || batch <= 0 || channels <= 0 || height <= 0 || width <= 0) return fail(FITV2_E_INVALID, "bad pack_uint8 argument");
    const int64_t n = (int64_t)batch * channels * height * width;
    CUDA_TRY(launch_k(pack_uint8_kernel, dim3(elementwise_grid(n * 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                      img, out, batch, channels, height, width));
    return FITV2_OK;
}

int fitv2_debug_gemm(fitv2_handle* h, int epilogue, const void* a, const void* w, const float* bias, float* out32, int M,
                     int N, int K, int bn, void* stream) {
    if (!h || !a || !w || !bias || !out32) return fail(FITV2_E_INVALID, "null argument");
    if (epilogue != EPI_PLAIN && epilogue != EPI_PLAIN + 1)
        return fail(FITV2_E_INVALID, "debug_gemm supports the plain epilogue only (3 = single CTA, 4 = 2-CTA multicast cluster)");
    const bool clustered = epilogue == EPI_PLAIN + 1;
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_map(&ma, a, h->cfg.operand_dtype, M, K, K, 128))) return rc;
    if ((rc = make_map(&mb, w, h->cfg.operand_dtype, N, K, K, clustered ? bn / 2 : bn))) return rc;
    GemmEpi ep;
    memset(&ep, 0, sizeof(ep));
    ep.bias = bias; ep.out32 = out32; ep.ld_out = N; ep.tokens = 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (clustered) {
        if (h->cfg.operand_dtype == FITV2_OPERAND_FP16) return launch_gemm_bn<EPI_PLAIN, __half, 2>(h, bn, ma, mb, M, N, K, 0, ep, st);
        return launch_gemm_bn<EPI_PLAIN, __nv_bfloat16, 2>(h, bn, ma, mb, M, N, K, 0, ep, st);
    }
    if (h->cfg.operand_dtype == FITV2_OPERAND_FP16) return launch_gemm_bn<EPI_PLAIN, __half, 1>(h, bn, ma, mb, M, N, K, 0, ep, st);
    return launch_gemm_bn<EPI_PLAIN, __nv_bfloat16, 1>(h, bn, ma, mb, M, N, K, 0, ep, st);
}

int fitv2_debug_attention(fitv2_handle* h, const void* q, const void* k, const void* vt, const float* mask, void* out,
                          int rows, int tokens, float* dbg_s, float* dbg_o, void* stream) {
    if (!h || !q || !k || !vt || !mask || !out) return fail(FITV2_E_INVALID, "null argument");
    if (!h->ws) return fail(FITV2_E_UNBOUND, "workspace is not set");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int* segu = reinterpret_cast<int*>(h->ws);            // first bytes of the workspace as scratch
    CUDA_TRY(launch_k(seg_uniform_kernel, dim3(rows), dim3(128), 0, st, 1, mask, segu, tokens));
    CUDA_TRY(cudaGetLastError());
    const int tokens_v = (tokens + 7) / 8 * 8;
    if (h->cfg.operand_dtype == FITV2_OPERAND_FP16)
        return launch_attention<__half>(h, q, k, vt, mask, segu, out, rows, tokens, tokens_v, st, dbg_s, dbg_o);
    return launch_attention<__nv_bfloat16>(h, q, k, vt, mask, segu, out, rows, tokens, tokens_v, st, dbg_s, dbg_o);
}

int fitv2_debug_tap(fitv2_handle* h, int what, void* dst, int64_t bytes, void* stream) {
    if (!h || !h->ws || !dst) return fail(FITV2_E_INVALID, "bad tap argument");
    const Layout& l = h->lay;
    if (!l.rows) return fail(FITV2_E_UNBOUND, "no forward has run yet");
    const size_t M = (size_t)l.rows * l.tokens, D = h->cfg.hidden_size;
    size_t off = 0, have = 0;
    switch (what) {
        case 0: off = l.c; have = (size_t)l.rows * D * 4; break;
        case 1: off = l.gmod; have = (size_t)l.rows * 6 * D * 4; break;
        case 2: off = l.mod; have = (size_t)h->cfg.depth * l.rows * 6 * D * 4; break;
        case 3: off = l.fmod; have = (size_t)l.rows * 2 * D * 4; break;
        case 4: off = l.x_res; have = M * D * 4; break;
        case 5: off = l.q; have = M * D * 2; break;
        case 6: off = l.k; have = M * D * 2; break;
        case 7: off = l.vt; have = (size_t)l.rows * D * l.tokens_v * 2; break;
        case 8: off = l.ao; have = M * D * 2; break;
        case 9: off = l.h; have = M * D * 2; break;
        case 10: off = l.hidden; have = M * h->cfg.mlp_hidden * 2; break;
        case 11: off = l.rope_cos; have = M * (h->cfg.head_dim / 2) * 4; break;
        case 12: off = l.rope_sin; have = M * (h->cfg.head_dim / 2) * 4; break;
        case 13: off = l.seg_uniform; have = (size_t)l.rows * 4; break;
        default: return fail(FITV2_E_INVALID, "unknown tap %d", what);
    }
    if ((size_t)bytes != have) return fail(FITV2_E_INVALID, "tap %d holds %lld bytes, caller asked for %lld", what, (long long)have, (long long)bytes);
    CUDA_TRY(cudaMemcpyAsync(dst, h->ws + off, have, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return FITV2_OK;
}

#ifdef FITV2_ATTN_TRACE
// timing experiments only (not part of include/fitv2_b200.h)
int fitv2_debug_attn_trace(unsigned long long* host_trace, unsigned int* host_n, int reset) {
    if (reset) {
        unsigned int z[20] = {0};
        CUDA_TRY(cudaMemcpyToSymbol(g_attn_trace_n, z, sizeof(z)));
        return FITV2_OK;
    }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(host_trace, g_attn_trace, sizeof(unsigned long long) * 20 * 512 * 2));
    CUDA_TRY(cudaMemcpyFromSymbol(host_n, g_attn_trace_n, sizeof(unsigned int) * 20));
    return FITV2_OK;
}
#endif

int fitv2_debug_layout(const fitv2_handle* h, int64_t* offsets, int64_t* sizes, int max_entries) {
    if (!h || !offsets || !sizes || max_entries <= 0) return fail(FITV2_E_INVALID, "bad debug_layout argument");
    if (!h->lay.rows) return fail(FITV2_E_UNBOUND, "no forward has run yet");
    const int n = h->lay.n_spans < max_entries ? h->lay.n_spans : max_entries;
    for (int i = 0; i < n; ++i) { offsets[i] = (int64_t)h->lay.span_off[i]; sizes[i] = (int64_t)h->lay.span_bytes[i]; }
    return n;
}

int64_t fitv2_kernel_launches(const fitv2_handle* h) { return h ? h->launches : 0; }

int fitv2_profile_set(fitv2_handle* h, uint32_t class_mask) {
    if (!h) return fail(FITV2_E_INVALID, "null handle");
    if (class_mask && h->prof_pool.empty()) {
        h->prof_pool.resize(8192);
        for (auto& e : h->prof_pool) {
            CUDA_TRY(cudaEventCreate(&e.a));
            CUDA_TRY(cudaEventCreate(&e.b));
        }
    }
    h->prof_mask = class_mask;
    h->prof_used = 0;
    h->prof_open = -1;
    return FITV2_OK;
}

int fitv2_profile_read(fitv2_handle* h, double* ms_sum, int64_t* count) {
    if (!h || !ms_sum || !count) return fail(FITV2_E_INVALID, "null argument");
    for (int i = 0; i < PC_COUNT; ++i) { ms_sum[i] = 0.0; count[i] = 0; }
    for (size_t i = 0; i < h->prof_used; ++i) {
        ProfEvt& e = h->prof_pool[i];
        CUDA_TRY(cudaEventSynchronize(e.b));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e.a, e.b));
        ms_sum[e.cls] += ms;
        count[e.cls]++;
    }
    h->prof_used = 0;
    return FITV2_OK;
}

}  // extern "C"
