// Memory-bound kernels of the FiTv2 hot path: 128-bit coalesced accesses, warp-shuffle reductions,
// fp32 math.  Reference lines are cited per kernel (paths relative to the reference repo).
#pragma once
#include "common.cuh"

namespace fitv2 {

// ---------------------------------------------------------------------------------------------
// CFG combine + Euler update  (sample_fitv2_ddp.py:310-314)
//   v = uncond + s * (cond - uncond) ;  z = z + dsigma * v        -- evaluated in exactly that order,
//   no FMA contraction, so the fp32 result is bit-identical to the PyTorch expression.
// z: (B, n) fp32 updated in place, v2: (2B, n): rows [0,B) cond, [B,2B) uncond.  n = tokens * C.
// ---------------------------------------------------------------------------------------------
__global__ void cfg_euler_kernel(float* __restrict__ z, const float* __restrict__ v2, float cfg_scale, float dsigma,
                                 const float* __restrict__ dsigma_dev, size_t half_elems)
{
    pdl_wait();
    pdl_launch_dependents();
    if (dsigma_dev) dsigma = *dsigma_dev;
    const size_t nvec = half_elems >> 2;
    const float4* cond = reinterpret_cast<const float4*>(v2);
    const float4* uncond = reinterpret_cast<const float4*>(v2 + half_elems);
    float4* zz = reinterpret_cast<float4*>(z);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const float4 c = cond[i], u = uncond[i];
        float4 o = zz[i];
        o.x = __fadd_rn(o.x, __fmul_rn(dsigma, __fadd_rn(u.x, __fmul_rn(cfg_scale, __fsub_rn(c.x, u.x)))));
        o.y = __fadd_rn(o.y, __fmul_rn(dsigma, __fadd_rn(u.y, __fmul_rn(cfg_scale, __fsub_rn(c.y, u.y)))));
        o.z = __fadd_rn(o.z, __fmul_rn(dsigma, __fadd_rn(u.z, __fmul_rn(cfg_scale, __fsub_rn(c.z, u.z)))));
        o.w = __fadd_rn(o.w, __fmul_rn(dsigma, __fadd_rn(u.w, __fmul_rn(cfg_scale, __fsub_rn(c.w, u.w)))));
        zz[i] = o;
    }
    // scalar tail (half_elems not a multiple of 4)
    for (size_t i = (nvec << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half_elems;
         i += (size_t)gridDim.x * blockDim.x) {
        const float c = v2[i], u = v2[half_elems + i];
        z[i] = __fadd_rn(z[i], __fmul_rn(dsigma, __fadd_rn(u, __fmul_rn(cfg_scale, __fsub_rn(c, u)))));
    }
}

// ---------------------------------------------------------------------------------------------
// Transport sampler updates (fit/scheduler/transport/integrators.py:29-48, transport.py:256-292, path.py:71-85),
// velocity model on the linear path.  All per-step scalars are computed on the host in fp32 with the reference's
// own expression order and read from a small device array (graph-replay friendly):
//   coef[0] rar      = alpha_t / d_alpha_t (= t)          coef[4] sqrt(2 * diffusion)
//   coef[1] var      = sigma^2 - rar * d_sigma * sigma    coef[5] sqrt(dt)
//   coef[2] diffusion(t)                                  coef[6] alpha_t          (Tweedie last step)
//   coef[3] dt  (or the last-step size)                   coef[7] sigma_t^2 / alpha_t
// Every product / sum below is a separate rounded fp32 operation in the order PyTorch evaluates the reference's
// expressions (no FMA contraction), so the results are bit-identical to it given the same velocity and noise.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sde_drift_elem(float x, float v, float rar, float var, float diff) {
    const float score = __fdiv_rn(__fsub_rn(__fmul_rn(rar, v), x), var);                // path.py:84
    return __fadd_rn(v, __fmul_rn(diff, score));                                        // transport.py:256-258
}

// Euler-Maruyama step, in place:  x <- (x + drift * dt) + sqrt(2 D) * (w * sqrt(dt))     (integrators.py:29-37)
// w == nullptr: the noise-free "Mean" last step  x <- x + drift * h                      (transport.py:277-280)
__global__ void sde_em_step_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ w,
                                   const float* __restrict__ coef, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    const float rar = coef[0], var = coef[1], diff = coef[2], dt = coef[3], s2d = coef[4], sdt = coef[5];
    const size_t nvec = n >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        float4 xv = reinterpret_cast<const float4*>(x)[i];
        const float4 vv = reinterpret_cast<const float4*>(v)[i];
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w) wv = reinterpret_cast<const float4*>(w)[i];
        float* xe = reinterpret_cast<float*>(&xv);
        const float* ve = reinterpret_cast<const float*>(&vv);
        const float* we = reinterpret_cast<const float*>(&wv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float mean = __fadd_rn(xe[j], __fmul_rn(sde_drift_elem(xe[j], ve[j], rar, var, diff), dt));
            xe[j] = w ? __fadd_rn(mean, __fmul_rn(s2d, __fmul_rn(we[j], sdt))) : mean;
        }
        reinterpret_cast<float4*>(x)[i] = xv;
    }
    for (size_t i = (nvec << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float mean = __fadd_rn(x[i], __fmul_rn(sde_drift_elem(x[i], v[i], rar, var, diff), dt));
        x[i] = w ? __fadd_rn(mean, __fmul_rn(s2d, __fmul_rn(w[i], sdt))) : mean;
    }
}

// out = v + D * score(v, x, t)   (the SDE drift; K1 / K2 of the Heun step, integrators.py:45-47)
__global__ void sde_drift_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ v,
                                 const float* __restrict__ coef, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    const float rar = coef[0], var = coef[1], diff = coef[2];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = sde_drift_elem(x[i], v[i], rar, var, diff);
}

// out = a + s2 * (s1 * b)      (xhat = x + sqrt(2D) * (w * sqrt(dt)); xp = xhat + dt * K1; ODE Euler y + dt * f; "Euler" last step)
__global__ void scaled_add_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
                                  const float* __restrict__ s, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    const float s1 = s[0], s2 = s[1];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __fadd_rn(a[i], __fmul_rn(s2, __fmul_rn(s1, b[i])));
}

// out = xhat + c * (k1 + k2),  c = 0.5 * dt                                              (integrators.py:48)
__global__ void heun_combine_kernel(float* __restrict__ out, const float* __restrict__ xhat, const float* __restrict__ k1,
                                    const float* __restrict__ k2, const float* __restrict__ c, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    const float cc = c[0];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __fadd_rn(xhat[i], __fmul_rn(cc, __fadd_rn(k1[i], k2[i])));
}

// "Tweedie" last step: out = x / alpha + (sigma^2 / alpha) * score(v, x, t)                (transport.py:281-286)
__global__ void tweedie_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ v,
                               const float* __restrict__ coef, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    const float rar = coef[0], var = coef[1], a = coef[6], c2 = coef[7];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float score = __fdiv_rn(__fsub_rn(__fmul_rn(rar, v[i]), x[i]), var);
        out[i] = __fadd_rn(__fdiv_rn(x[i], a), __fmul_rn(c2, score));
    }
}

// ---------------------------------------------------------------------------------------------
// After the trajectory (sample_fitv2_ddp.py:319-324):
//   latents = unpatchify(z) / scaling_factor    (fit_model.py:171-187: "b (h w) (c p1 p2) -> b c (h p1) (w p2)")
//   ... vae.decode (reference PyTorch, outside this path) ...
//   images  = clamp(127.5 * clamp(s, -1, 1) + 128, 0, 255).permute(0, 2, 3, 1).to(uint8)
// Both are pure index permutations with one correctly rounded fp32 operation per element (IEEE division; multiply
// then add, no FMA; float -> uint8 truncation), so they are bit-identical to the PyTorch expressions.
// ---------------------------------------------------------------------------------------------
// z: (B, hp*wp, C*p*p) fp32 -> out: (B, C, hp*p, wp*p) fp32.  One thread per output element: consecutive threads write
// consecutive W positions; the matching reads are p-strided inside one token row (64 bytes for p = 2, C = 4).
__global__ void unpatchify_scale_kernel(const float* __restrict__ z, float* __restrict__ out, float scaling_factor,
                                        int B, int hp, int wp, int C, int p)
{
    pdl_wait();
    pdl_launch_dependents();
    const int H = hp * p, W = wp * p;
    const size_t total = (size_t)B * C * H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((size_t)W * H)) % C), b = (int)(i / ((size_t)W * H * C));
        const int token = (y / p) * wp + (x / p);
        const int ch = (c * p + (y % p)) * p + (x % p);
        const float v = z[((size_t)b * hp * wp + token) * (C * p * p) + ch];
        out[i] = scaling_factor == 1.0f ? v : __fdiv_rn(v, scaling_factor);
    }
}

// img: (B, C, H, W) fp32 -> out: (B, H, W, C) uint8
__global__ void pack_uint8_kernel(const float* __restrict__ img, unsigned char* __restrict__ out, int B, int C, int H, int W)
{
    pdl_wait();
    pdl_launch_dependents();
    const size_t total = (size_t)B * C * H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C), x = (int)((i / C) % W), y = (int)((i / ((size_t)C * W)) % H), b = (int)(i / ((size_t)C * W * H));
        float s = img[(((size_t)b * C + c) * H + y) * W + x];
        s = fminf(fmaxf(s, -1.0f), 1.0f);                                               // samples.clamp(-1, 1)
        s = __fadd_rn(__fmul_rn(127.5f, s), 128.0f);
        s = fminf(fmaxf(s, 0.0f), 255.0f);                                              // torch.clamp(..., 0, 255)
        out[i] = (unsigned char)(int)s;                                                 // .to(torch.uint8): truncation
    }
}

// Fixed-grid Runge-Kutta stages of torchdiffeq.odeint (the reference's sample_ode route, integrators.py:109-116; torchdiffeq
// itself is an un-vendored dependency: the expressions below restate its rk_common.py / fixed_grid.py step functions).
// Every operation is a separately rounded fp32 operation in PyTorch's evaluation order.  s[0] = dt (0-dim tensor in the
// reference), s[1..3] = the Python-float coefficients rounded to fp32.
//   mode 0: out = y + (dt * k1) * s1                                 (rk2 / rk3 / rk4 second-stage argument:  y0 + dt * k1 * a21)
//   mode 1: out = y + dt * (k1 * s1 + k2 * s2)                       (rk3 third-stage argument; rk2 final update)
//   mode 2: out = y + dt * (k1 * s1 + k2 * s2 + k3 * s3)             (rk3 final update)
//   mode 3: out = y + dt * (k2 - k1 * s1)                            (rk4 3/8-rule third-stage argument)
//   mode 4: out = y + dt * (k1 - k2 + k3)                            (rk4 fourth-stage argument)
//   mode 5: out = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125         (rk4 3/8-rule final update)
__global__ void rk_stage_kernel(float* __restrict__ out, const float* __restrict__ y, const float* __restrict__ k1,
                                const float* __restrict__ k2, const float* __restrict__ k3, const float* __restrict__ k4,
                                const float* __restrict__ s, int mode, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    const float dt = s[0], s1 = s[1], s2 = s[2], s3 = s[3];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float d;
        switch (mode) {
            case 0: d = __fmul_rn(__fmul_rn(dt, k1[i]), s1); break;
            case 1: d = __fmul_rn(dt, __fadd_rn(__fmul_rn(k1[i], s1), __fmul_rn(k2[i], s2))); break;
            case 2: d = __fmul_rn(dt, __fadd_rn(__fadd_rn(__fmul_rn(k1[i], s1), __fmul_rn(k2[i], s2)), __fmul_rn(k3[i], s3))); break;
            case 3: d = __fmul_rn(dt, __fsub_rn(k2[i], __fmul_rn(k1[i], s1))); break;
            case 4: d = __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[i], k2[i]), k3[i])); break;
            default: d = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(k1[i], __fmul_rn(3.0f, __fadd_rn(k2[i], k3[i]))), k4[i]), dt), 0.125f); break;
        }
        out[i] = __fadd_rn(y[i], d);
    }
}

// Adaptive Dormand-Prince 5(4) ("dopri5", the reference's default --ode-sampling-method through torchdiffeq.odeint): the two
// device-side pieces of a step.
//   lincomb:     out = c[0] * y + sum_{i < nk} c[1 + i] * k_i        (stage arguments, the 5th-order solution, the error estimate,
//                                                                     the dense-output polynomial and its evaluation)
//   scaled_rms:  out[0] = sqrt(mean(((a - b) / (atol + rtol * |s|))^2))   (b, s nullable: plain RMS norm, torchdiffeq's default `norm`)
// One block, fixed summation order (deterministic); the state is at most a few 10^5 elements and the solver synchronises on
// this scalar anyway to accept / reject the step.
struct LinComb { const float* k[7]; };
__global__ void lincomb_kernel(float* __restrict__ out, const float* __restrict__ y, LinComb ks, const float* __restrict__ c, int nk, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    float cc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cc[i] = i <= nk ? c[i] : 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = __fmul_rn(cc[0], y[i]);
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (j < nk) acc = __fadd_rn(acc, __fmul_rn(cc[1 + j], ks.k[j][i]));
        out[i] = acc;
    }
}

// One thread-block CLUSTER of kRmsCluster CTAs (grid = one cluster): every CTA reduces a contiguous slice, the partial sums meet in
// CTA 0's shared memory through DSMEM and are added in rank order, so the result does not depend on scheduling (the dopri5
// controller branches on it).  A single CTA was issue-bound on one SM: 67 us at 131 k elements (ncu), 31 us with four element
// streams per thread.
constexpr int kRmsCluster = 8;
__global__ void __launch_bounds__(1024)
scaled_rms_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ s,
                  float atol, float rtol, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    __shared__ double red[32];
    __shared__ double part[kRmsCluster];
    const uint32_t rank = cluster_ctarank();
    const size_t per = (n + kRmsCluster - 1) / kRmsCluster;
    const size_t lo = rank * per, hi = lo + per < n ? lo + per : n;
    auto term = [&](size_t i) -> double {
        float v = b ? __fsub_rn(a[i], b[i]) : a[i];
        if (s) v = __fdiv_rn(v, __fadd_rn(atol, __fmul_rn(rtol, fabsf(s[i]))));
        return (double)v * (double)v;
    };
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};
    const size_t B = blockDim.x;
    size_t i = lo + threadIdx.x;
    for (; i + 3 * B < hi; i += 4 * B) {
        const double t0 = term(i), t1 = term(i + B), t2 = term(i + 2 * B), t3 = term(i + 3 * B);
        acc4[0] += t0; acc4[1] += t1; acc4[2] += t2; acc4[3] += t3;
    }
    for (; i < hi; i += B) acc4[0] += term(i);
    double acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        // part[rank] of CTA 0 (shared::cluster address of the same variable in the CTA of rank 0)
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&part[rank])), "r"(0u));
        asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(remote), "d"(t) : "memory");
    }
    cluster_sync();                                                     // release / acquire: the eight stores are visible to CTA 0
    if (rank == 0 && threadIdx.x == 0) {
        double t = 0.0;
        for (int r = 0; r < kRmsCluster; ++r) t += part[r];
        out[0] = (float)sqrt(t / (double)n);
    }
}

// (rows, A, B) -> (rows, B, A): the 'B C N' <-> 'B N C' rearranges of the use_sit = False layout (fit_model.py:204,231).
__global__ void transpose_inner_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int A, int B)
{
    pdl_wait();
    pdl_launch_dependents();
    const size_t total = (size_t)rows * A * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(i % A), b = (int)((i / A) % B);
        const size_t r = i / ((size_t)A * B);
        out[i] = in[(r * A + a) * B + b];
    }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n)
{
    pdl_wait();
    pdl_launch_dependents();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __float2half_rn(in[i]);
}

// forward_with_cfg channel-limited guidance (fit_model.py:253-275): channels [0, c_cfg) of BOTH halves become
// uncond + s_b * (cond - uncond); channels >= c_cfg pass through.  out: (2B, tokens, C) in place.
// scale_per_sample may be null (then `scale` is used for every sample).
__global__ void cfg_combine_kernel(float* __restrict__ out, const float* __restrict__ scale_per_sample, float scale,
                                   int B, int tokens, int C, int c_cfg)
{
    pdl_wait();
    pdl_launch_dependents();
    const size_t per_sample = (size_t)tokens * C;
    const size_t total = (size_t)B * per_sample;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C);
        if (ch >= c_cfg) continue;
        const float s = scale_per_sample ? scale_per_sample[i / per_sample] : scale;
        const float c = out[i], u = out[total + i];
        const float g = __fadd_rn(u, __fmul_rn(s, __fsub_rn(c, u)));
        out[i] = g;
        out[total + i] = g;
    }
}

// ---------------------------------------------------------------------------------------------
// Patch embedding  (fit/model/modules.py:34-37): x[m, :] = W (D x Cin) * xin[m % rows_in, :] + b
// K = 16 -> bandwidth-bound (writes M*D fp32).  One block = kPatchRows token rows; a thread owns 4 consecutive
// output features and keeps their weight rows in registers (64 weight loads amortised over kPatchRows * 4 outputs; with 8
// rows per block every block re-read the whole weight matrix for 32 outputs per thread: 59 us against 12 us of writes).
// Launched with D / 4 threads (rounded up to a warp) so that one pass covers the row.
// ---------------------------------------------------------------------------------------------
constexpr int kPatchRows = 8;
// Persistent blocks: a thread loads its 4 weight rows ONCE and then walks groups of kPatchRows token rows (grid-stride).
// The kernel is ISSUE-bound, not write-bound (a plain 75 MB fill runs at 5.6 TB/s = 14 us on this GPU; the scalar version took 36 us
// for its 85 instructions per 16 output bytes): the 64 FMAs per float4 are 32 packed `fma.rn.f32x2` (two independent IEEE FMAs per
// instruction, so the result is bit-identical to the scalar chain), the (x, x) operand pairs are pre-packed in shared memory, and the
// next group's inputs are fetched into registers before the current group is computed.
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int CIN>
__global__ void __launch_bounds__(576)
patch_embed_kernel(const float* __restrict__ xin, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ x, int M, int D, int rows_in_tokens /* tokens * rows_in */)
{
    pdl_wait();
    pdl_launch_dependents();
    __shared__ __align__(16) uint64_t sx[2][kPatchRows][CIN];           // (x, x) pairs
    constexpr int kGroupElems = kPatchRows * CIN, kPref = 4;            // up to kPref input elements per thread and group (>= 32 threads)
    const int groups = (M + kPatchRows - 1) / kPatchRows;
    const int d0 = threadIdx.x * 4;                                     // launched with >= D / 4 threads (D <= 2304): one pass covers the row
    const bool active = d0 < D;
    uint64_t wr[2][CIN];                                                // (w[d0][c], w[d0+1][c]) and (w[d0+2][c], w[d0+3][c])
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
#pragma unroll
        for (int c = 0; c < CIN; c += 4) {
            float4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = __ldg(reinterpret_cast<const float4*>(w + (size_t)(d0 + j) * CIN + c));
            wr[0][c] = pack_f32x2(t[0].x, t[1].x); wr[0][c + 1] = pack_f32x2(t[0].y, t[1].y);
            wr[0][c + 2] = pack_f32x2(t[0].z, t[1].z); wr[0][c + 3] = pack_f32x2(t[0].w, t[1].w);
            wr[1][c] = pack_f32x2(t[2].x, t[3].x); wr[1][c + 1] = pack_f32x2(t[2].y, t[3].y);
            wr[1][c + 2] = pack_f32x2(t[2].z, t[3].z); wr[1][c + 3] = pack_f32x2(t[2].w, t[3].w);
        }
        bb = __ldg(reinterpret_cast<const float4*>(b + d0));
    } else {
#pragma unroll
        for (int c = 0; c < CIN; ++c) wr[0][c] = wr[1][c] = 0ull;
    }
    float pre[kPref];
    auto fetch = [&](int g) {
#pragma unroll
        for (int k = 0; k < kPref; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            const int m = g * kPatchRows + i / CIN;
            pre[k] = (g < groups && i < kGroupElems && m < M) ? __ldg(xin + (size_t)(m % rows_in_tokens) * CIN + (i % CIN)) : 0.f;
        }
    };
    fetch(blockIdx.x);
    int buf = 0;
    for (int g = blockIdx.x; g < groups; g += gridDim.x, buf ^= 1) {
        const int m0 = g * kPatchRows;
#pragma unroll
        for (int k = 0; k < kPref; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < kGroupElems) sx[buf][i / CIN][i % CIN] = pack_f32x2(pre[k], pre[k]);
        }
        __syncthreads();                                                // (double-buffered: the next group's fill does not race this group's reads)
        fetch(g + gridDim.x);
        if (active) {
            // two rows at a time: four independent FMA chains per thread, the (x, x) pairs fetched two at a time (128-bit shared loads)
#pragma unroll
            for (int r = 0; r < kPatchRows; r += 2) {
                if (m0 + r >= M) break;
                uint64_t acc[2][2] = {{0ull, 0ull}, {0ull, 0ull}};
#pragma unroll
                for (int c = 0; c < CIN; c += 2) {
                    const ulonglong2 xa = *reinterpret_cast<const ulonglong2*>(&sx[buf][r][c]);
                    const ulonglong2 xb = *reinterpret_cast<const ulonglong2*>(&sx[buf][r + 1][c]);
                    acc[0][0] = fma_f32x2(xa.x, wr[0][c], acc[0][0]);
                    acc[0][1] = fma_f32x2(xa.x, wr[1][c], acc[0][1]);
                    acc[1][0] = fma_f32x2(xb.x, wr[0][c], acc[1][0]);
                    acc[1][1] = fma_f32x2(xb.x, wr[1][c], acc[1][1]);
                    acc[0][0] = fma_f32x2(xa.y, wr[0][c + 1], acc[0][0]);
                    acc[0][1] = fma_f32x2(xa.y, wr[1][c + 1], acc[0][1]);
                    acc[1][0] = fma_f32x2(xb.y, wr[0][c + 1], acc[1][0]);
                    acc[1][1] = fma_f32x2(xb.y, wr[1][c + 1], acc[1][1]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (m0 + r + q < M) {
                        float a0, a1, a2, a3;
                        unpack_f32x2(acc[q][0], a0, a1);
                        unpack_f32x2(acc[q][1], a2, a3);
                        *reinterpret_cast<float4*>(x + (size_t)(m0 + r + q) * D + d0) = make_float4(a0 + bb.x, a1 + bb.y, a2 + bb.z, a3 + bb.w);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (no affine, eps 1e-6) + adaLN modulate -> 16-bit GEMM operand
//   h = LN(x) * (1 + scale[sample]) + shift[sample]      (norms.py:41-42, fit/model/utils.py:6-7, modules.py:272-273)
// One warp per token row, the row lives in registers (NV float4 per lane), two-pass fp32 statistics.
// ---------------------------------------------------------------------------------------------
// MODE 0: affine-free LayerNorm (the FiTv2 configs); 1: LayerNorm * weight ('w_layernorm', norms.py:35-38);
// 2: RMSNorm * weight ('rmsnorm', norms.py:53-77).  The weight vector (D floats, cache-resident) is fetched in the output loop.
template <typename OT, int NV, int MODE = 0, bool EXACT = false>   // EXACT: D == NV * 128, no tail guards
__global__ void __launch_bounds__(256)
ln_modulate_kernel(const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
                   int mod_ld, OT* __restrict__ h, int M, int D, int tokens, const float* __restrict__ norm_w)
{
    pdl_wait();
    pdl_launch_dependents();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= M) return;
    const int m = warp;
    const int nvec = D >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)m * D);
    const unsigned sample = (unsigned)m / (unsigned)tokens;
    const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)sample * mod_ld);
    const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)sample * mod_ld);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // every load of the row (x, shift, scale) is issued before the first use: 3*NV 128-bit requests in flight per lane
    float4 v[NV], a[NV], g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = lane + 32 * i;
        v[i] = (EXACT || j < nvec) ? ld_stream_f4(xr + j) : zero4;
    }
    // wider rows (3B) would not fit the register file.  (Measured alternative: fetching shift / scale late, 64 registers and
    // twice the resident warps, is SLOWER: 35 vs 28 us per launch.)
    constexpr bool kEarlyMod = NV <= 9;
    if constexpr (kEarlyMod) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = lane + 32 * i;
            a[i] = (EXACT || j < nvec) ? __ldg(sh + j) : zero4;
            g[i] = (EXACT || j < nvec) ? __ldg(sc + j) : zero4;
        }
    }
    const float inv_d = 1.0f / (float)D;
    float mean = 0.f;
    if constexpr (MODE != 2) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        mean = warp_sum(s) * inv_d;
    }
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = lane + 32 * i;
        if (EXACT || j < nvec) {
            const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
            q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + 1e-6f);
    uint2* hr = reinterpret_cast<uint2*>(h + (size_t)m * D);
    const float4* nw = reinterpret_cast<const float4*>(norm_w);
    // The kernel is issue-bound as much as memory-bound (ncu: 712 SASS instructions per row, half of the launch in issue slots), so
    // the output is three FMAs per element: (v - mean) * rstd * (1 + g) + a = v * A + B with A = rstd * (1 + g) [* w], B = a - mean * A
    const float nmean = -mean;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = lane + 32 * i;
        if (EXACT || j < nvec) {
            if constexpr (!kEarlyMod) { a[i] = __ldg(sh + j); g[i] = __ldg(sc + j); }
            float A0 = fmaf(rstd, g[i].x, rstd), A1 = fmaf(rstd, g[i].y, rstd), A2 = fmaf(rstd, g[i].z, rstd), A3 = fmaf(rstd, g[i].w, rstd);
            if constexpr (MODE != 0) { const float4 w4 = __ldg(nw + j); A0 *= w4.x; A1 *= w4.y; A2 *= w4.z; A3 *= w4.w; }
            const float o0 = fmaf(v[i].x, A0, fmaf(nmean, A0, a[i].x));
            const float o1 = fmaf(v[i].y, A1, fmaf(nmean, A1, a[i].y));
            const float o2 = fmaf(v[i].z, A2, fmaf(nmean, A2, a[i].z));
            const float o3 = fmaf(v[i].w, A3, fmaf(nmean, A3, a[i].w));
            hr[j] = make_uint2(Op16<OT>::pack(o0, o1), Op16<OT>::pack(o2, o3));
        }
    }
}

// Wide rows (3B/2: D = 2304): one row per WARP PAIR, so that the row, its shift and its scale all fit in registers and every
// load is issued up front like in the narrow kernel (one warp per 9 KB row with late modulation loads reached 2.0 TB/s).
// The two partial statistics meet in shared memory.  256 threads = 4 rows per CTA.
template <typename OT, int NV>
__global__ void __launch_bounds__(256)
ln_modulate_pair_kernel(const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
                        int mod_ld, OT* __restrict__ h, int M, int D, int tokens)
{
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[2][8];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * 4 + (wib >> 1);
    const int t = (wib & 1) * 32 + lane;                 // 0..63 inside the pair
    const int nvec = D >> 2;
    const bool row_ok = m < M;
    const int mm = row_ok ? m : M - 1;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)mm * D);
    const int sample = mm / tokens;
    const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)sample * mod_ld);
    const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)sample * mod_ld);
    float4 v[NV], a[NV], g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = t + 64 * i;
        v[i] = (j < nvec) ? ld_stream_f4(xr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = t + 64 * i;
        a[i] = (j < nvec) ? __ldg(sh + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        g[i] = (j < nvec) ? __ldg(sc + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    s = warp_sum(s);
    if (lane == 0) red[0][wib] = s;
    __syncthreads();
    const float mean = (red[0][wib & ~1] + red[0][wib | 1]) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (t + 64 * i < nvec) {
            const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
            q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
    }
    q = warp_sum(q);
    if (lane == 0) red[1][wib] = q;
    __syncthreads();
    const float rstd = rsqrtf((red[1][wib & ~1] + red[1][wib | 1]) / (float)D + 1e-6f);
    if (!row_ok) return;
    uint2* hr = reinterpret_cast<uint2*>(h + (size_t)m * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = t + 64 * i;
        if (j < nvec) {
            const float A0 = fmaf(rstd, g[i].x, rstd), A1 = fmaf(rstd, g[i].y, rstd), A2 = fmaf(rstd, g[i].z, rstd), A3 = fmaf(rstd, g[i].w, rstd);
            const float o0 = fmaf(v[i].x, A0, fmaf(-mean, A0, a[i].x));     // = (v - mean) * rstd * (1 + g) + a, three FMAs per element
            const float o1 = fmaf(v[i].y, A1, fmaf(-mean, A1, a[i].y));
            const float o2 = fmaf(v[i].z, A2, fmaf(-mean, A2, a[i].z));
            const float o3 = fmaf(v[i].w, A3, fmaf(-mean, A3, a[i].w));
            hr[j] = make_uint2(Op16<OT>::pack(o0, o1), Op16<OT>::pack(o2, o3));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Final layer  (modules.py:292-296, fit_model.py:230):
//   out[m, :] = (W_o (Cout x D) * (LN(x[m]) * (1 + scale) + shift) + b_o) * mask[m]
// N = 16 -> bandwidth-bound (reads M*D fp32).  One warp per row, W_o cached in shared memory.
// ---------------------------------------------------------------------------------------------
template <int NV, int COUT>
__global__ void __launch_bounds__(256)
final_layer_kernel(const float* __restrict__ x, const float* __restrict__ fmod /* (samples, 2D): shift | scale */,
                   const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ mask,
                   float* __restrict__ out, int M, int D, int tokens,
                   const float* __restrict__ norm_w /* null, or (D) weight of norm_final */, int rms /* 1: RMSNorm instead of LayerNorm */,
                   int out_ld /* output channels per token (16, or 32 with learn_sigma) */, int out_col0 /* first channel of this launch */)
{
    pdl_wait();
    pdl_launch_dependents();
    w += (size_t)out_col0 * D;
    b += out_col0;
    extern __shared__ float sw[];                       // COUT * D
    {   // 73 KB of weights per block: 128-bit loads, several in flight per thread
        const float4* w4 = reinterpret_cast<const float4*>(w);
        float4* sw4 = reinterpret_cast<float4*>(sw);
        const int n4 = COUT * D / 4;
#pragma unroll 6
        for (int i = threadIdx.x; i < n4; i += blockDim.x) sw4[i] = __ldg(w4 + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int nvec = D >> 2;
    // Two rows per warp and iteration: every weight vector read from shared memory serves both rows (the kernel was bound by
    // the 73 KB of shared-memory reads per row, 87 us for 75 MB of x), and the second row's loads overlap the first row's math.
    constexpr int RW = 2;
    const int stride = gridDim.x * warps_per_block;
    for (int m0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5); m0 < M; m0 += RW * stride) {
        float4 v[RW][NV];
        bool ok[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const int m = m0 + r * stride;
            ok[r] = m < M;
            const float4* xr = reinterpret_cast<const float4*>(x + (size_t)(ok[r] ? m : m0) * D);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int j = lane + 32 * i;
                v[r][i] = (j < nvec) ? xr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const int m = ok[r] ? m0 + r * stride : m0;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
            const float mean = rms ? 0.f : warp_sum(s) / (float)D;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int j = lane + 32 * i;
                if (j < nvec) {
                    const float a = v[r][i].x - mean, bq = v[r][i].y - mean, c = v[r][i].z - mean, d = v[r][i].w - mean;
                    q += (a * a + bq * bq) + (c * c + d * d);
                }
            }
            const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-6f);
            const int sample = m / tokens;
            const float4* sh = reinterpret_cast<const float4*>(fmod + (size_t)sample * 2 * D);
            const float4* sc = reinterpret_cast<const float4*>(fmod + (size_t)sample * 2 * D + D);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int j = lane + 32 * i;
                if (j < nvec) {
                    const float4 a = __ldg(sh + j), g = __ldg(sc + j);
                    float4 nw = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (norm_w) nw = __ldg(reinterpret_cast<const float4*>(norm_w) + j);
                    v[r][i].x = (v[r][i].x - mean) * rstd * nw.x * (1.f + g.x) + a.x;
                    v[r][i].y = (v[r][i].y - mean) * rstd * nw.y * (1.f + g.y) + a.y;
                    v[r][i].z = (v[r][i].z - mean) * rstd * nw.z * (1.f + g.z) + a.z;
                    v[r][i].w = (v[r][i].w - mean) * rstd * nw.w * (1.f + g.w) + a.w;
                }
            }
        }
        float acc[RW][COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            const float4* wr = reinterpret_cast<const float4*>(sw + (size_t)o * D);
            float a[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r) a[r] = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int j = lane + 32 * i;
                if (j < nvec) {
                    const float4 ww = wr[j];
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        a[r] = fmaf(v[r][i].x, ww.x, a[r]); a[r] = fmaf(v[r][i].y, ww.y, a[r]);
                        a[r] = fmaf(v[r][i].z, ww.z, a[r]); a[r] = fmaf(v[r][i].w, ww.w, a[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RW; ++r) acc[r][o] = a[r];
        }
        // 16 warp reductions at once: a halving butterfly (8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5); afterwards
        // lanes 2k and 2k+1 hold the total of output  o = bits (4,3,2,1) of the lane, most significant first
        static_assert(COUT == 16, "the reduction butterfly is written for 16 outputs");
#pragma unroll
        for (int r = 0; r < RW; ++r) {
#pragma unroll
            for (int step = 0; step < 4; ++step) {
                const int off = 16 >> step, half = 8 >> step;         // lane bit `off` selects which half of the values a lane keeps
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int o = 0; o < half; ++o) {
                    const float send = up ? acc[r][o] : acc[r][o + half];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                    acc[r][o] = (up ? acc[r][o + half] : acc[r][o]) + recv;
                }
            }
            const float total = acc[r][0] + __shfl_xor_sync(0xffffffffu, acc[r][0], 1);
            const int m = m0 + r * stride;
            if (ok[r] && (lane & 1) == 0) {
                const int o = lane >> 1;                               // bits (4,3,2,1) -> output index
                out[(size_t)m * out_ld + out_col0 + o] = (total + __ldg(b + o)) * mask[m];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RoPE tables  (rope.py:165-170, 308-333): for token row m of sample s
//   pair i <  dh/4 : angle = float(grid[s,1,n]) * freqs_h[i]            (h position)
//   pair i >= dh/4 : angle = float(grid[s,0,n]) * freqs_w[i - dh/4]     (w position)
//   cos/sin tables (dh/2, M) fp32 (pair-major), optionally scaled by the yarn / ntk-pro magnitude.
// ---------------------------------------------------------------------------------------------
// freq_stride = 0: one frequency vector per axis for the whole batch (cached mode); freq_stride = dh/4: per-sample
// vectors (online mode, rope.py:234-274: dynamic NTK scale from each sample's own size).
__global__ void rope_table_kernel(const long long* __restrict__ grid, const float* __restrict__ freqs_h,
                                  const float* __restrict__ freqs_w, int freq_stride, float mag, float* __restrict__ cos_t,
                                  float* __restrict__ sin_t, int samples, int tokens, int half /* dh/2 */)
{
    pdl_wait();
    pdl_launch_dependents();
    const int quarter = half >> 1;
    const size_t total = (size_t)samples * tokens * half;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t M = (size_t)samples * tokens;           // pair-major layout [dh/2][M]: lanes -> consecutive token rows
        const int p = (int)(i / M);
        const size_t m = i % M;
        const int n = (int)(m % tokens);
        const int s = (int)(m / tokens);
        const long long* g = grid + (size_t)s * 2 * tokens;
        float ang;
        if (p < quarter) ang = __fmul_rn((float)g[tokens + n], freqs_h[(size_t)s * freq_stride + p]);
        else             ang = __fmul_rn((float)g[n], freqs_w[(size_t)s * freq_stride + p - quarter]);
        float sv, cv;
        sincosf(ang, &sv, &cv);
        if (mag != 1.0f) { cv = __fmul_rn(cv, mag); sv = __fmul_rn(sv, mag); }
        cos_t[i] = cv;
        sin_t[i] = sv;
    }
}

// Per-sample KEY LENGTH for the attention kernels (they then skip the per-element segment compares):
//   tokens      every segment id of the sample is identical (the sampling scripts: mask of ones);
//   1..tokens-1 the ids are one non-zero value followed by zeros only (a padded sample of a mixed-aspect batch: n valid tokens, then
//               padding).  Valid queries see exactly the first n keys; padded queries (id 0) may see anything, their output rows
//               are zeroed (`* (mask != 0)`, fit/model/modules.py:204);
//   0           anything else (packed sequences with several ids): the kernels compare ids per element.
__global__ void seg_uniform_kernel(const float* __restrict__ seg, int* __restrict__ klen, int tokens)
{
    pdl_wait();
    pdl_launch_dependents();
    const float* s = seg + (size_t)blockIdx.x * tokens;
    const float first = s[0];
    int n_first = 0, prefix_ok = 1;                                     // ids equal to the first one: contiguous from the start; the rest is 0
    for (int i = threadIdx.x; i < tokens; i += blockDim.x) {
        const float v = s[i];
        if (v == first) { ++n_first; prefix_ok &= (i == 0 || s[i - 1] == first); }
        else prefix_ok &= (v == 0.f);
    }
    __shared__ int cnt[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_first += __shfl_xor_sync(0xffffffffu, n_first, o);
    if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = n_first;
    prefix_ok = __syncthreads_and(prefix_ok);
    if (threadIdx.x == 0) {
        int n = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) n += cnt[w];
        klen[blockIdx.x] = prefix_ok ? n : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Timestep features  (fit_model.py:202-203, modules.py:52-71):
//   t' = min(ts*t / (1 + (ts-1)*t), 1);  te = [cos(t' f_i), sin(t' f_i)], f_i = exp(-ln(1e4) * i / 128)
// ---------------------------------------------------------------------------------------------
__global__ void timestep_features_kernel(const float* __restrict__ t, float time_shifting, float* __restrict__ te,
                                         int samples)
{
    pdl_wait();
    pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= samples * 128) return;
    const int s = i >> 7, j = i & 127;
    float tt = t[s];
    tt = __fdiv_rn(__fmul_rn(time_shifting, tt), __fadd_rn(1.0f, __fmul_rn(time_shifting - 1.0f, tt)));
    tt = fminf(tt, 1.0f);
    const float f = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, (float)j), 128.0f));
    const float a = __fmul_rn(tt, f);
    te[(size_t)s * 256 + j] = cosf(a);
    te[(size_t)s * 256 + 128 + j] = sinf(a);
}

// ---------------------------------------------------------------------------------------------
// Small-M fp32 linear for the conditioning path (t-MLP, global adaLN, adaLN-LoRA, final adaLN):
//   out[z][r, n] = sum_k act(A[z][r, k]) * W[z][n, k] + bias[z][n] (+ add[r, n]) (+ emb[label[r], n])
// M = samples (64 with CFG) -> weight-bandwidth-bound; fp32 FMA keeps the modulation exact to
// fp32 rounding because shift/scale/gate feed every token of every block.
// 64x64 output tile per CTA, 16x16 threads with 4x4 register tiles, K chunks of 16 through smem.
// ---------------------------------------------------------------------------------------------
struct SmallLinear {
    const float* A; size_t a_batch_stride; int lda;
    const float* W; size_t w_batch_stride;
    const float* bias; size_t bias_batch_stride;
    const float* add;            // (rows, N) or null, shared by all batches
    const float* mul; size_t mul_batch_stride; int ldm;   // optional: out = (acc + bias) * mul[batch][r][n]  (SwiGLU gate, adaln_type 'swiglu')
    const float* emb; const long long* labels;   // optional embedding-row add (label gather), ld = N
    int num_emb; int* err;       // rows of the table; labels outside [0, num_emb) set bit 0 of *err and read row 0 (the reference raises)
    float* out; size_t out_batch_stride; int ldo;
    float* out_silu;             // optional second output: silu(out) (same layout)
    float* out_silu_split;       // optional: tf32 hi / lo split of silu(out), stacked [rows/64][128][N] (operand of cond_tc.cuh)
    int rows, N, K;
    int act_silu_in;             // apply SiLU to A on load
    int ksplit;                  // > 1: K is split over blockIdx.y; raw partial sums go to `partial`, finalize adds the rest
    float* partial;              // [batches][ksplit][rows][N]
};

__global__ void __launch_bounds__(256)
small_linear_kernel(SmallLinear p)
{
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int z = blockIdx.z;
    const float* A = p.A + z * p.a_batch_stride;
    const float* W = p.W + z * p.w_batch_stride;
    const int ks = p.ksplit > 1 ? blockIdx.y % p.ksplit : 0;
    const int n0 = blockIdx.x * 64, r0 = (p.ksplit > 1 ? blockIdx.y / p.ksplit : blockIdx.y) * 64;
    // launches with few output tiles (N = 1152 .. 2304 at 64 rows) would run on a handful of SMs, each walking the whole
    // K extent serially (latency bound): split K so that every launch fills the GPU; the split sums are added in a
    // fixed order by small_linear_finalize_kernel (deterministic, no atomics).
    const int kchunk = p.ksplit > 1 ? ((p.K + 16 * p.ksplit - 1) / (16 * p.ksplit)) * 16 : p.K;
    const int k_begin = ks * kchunk, k_end = min(p.K, k_begin + kchunk);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lr = threadIdx.x >> 2;            // 0..63 : row (A) / col (W) loaded by this thread
    const int lk = (threadIdx.x & 3) * 4;       // 0,4,8,12
    // register double buffering: the global loads of chunk k + 1 are in flight while chunk k is multiplied out of shared memory
    auto load_chunk = [&](int k0, float4& av, float4& wv) {
        av = make_float4(0.f, 0.f, 0.f, 0.f); wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + lr < p.rows && k0 + lk < k_end) {
            av = *reinterpret_cast<const float4*>(A + (size_t)(r0 + lr) * p.lda + k0 + lk);
            if (p.act_silu_in) {
                av.x = av.x / (1.f + expf(-av.x)); av.y = av.y / (1.f + expf(-av.y));
                av.z = av.z / (1.f + expf(-av.z)); av.w = av.w / (1.f + expf(-av.w));
            }
        }
        if (n0 + lr < p.N && k0 + lk < k_end)
            wv = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * p.K + k0 + lk));
    };
    float4 av, wv;
    if (k_begin < k_end) load_chunk(k_begin, av, wv);
    for (int k0 = k_begin; k0 < k_end; k0 += 16) {
        __syncthreads();
        As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
        Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
        __syncthreads();
        if (k0 + 16 < k_end) load_chunk(k0 + 16, av, wv);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
    }
    if (p.ksplit > 1) {
        float* part = p.partial + ((size_t)z * p.ksplit + ks) * p.rows * p.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty * 4 + i;
            if (r >= p.rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n < p.N) part[(size_t)r * p.N + n] = acc[i][j];
            }
        }
        return;
    }
    const float* bias = p.bias ? p.bias + z * p.bias_batch_stride : nullptr;
    float* out = p.out + z * p.out_batch_stride;
    float* out_silu = p.out_silu ? p.out_silu + z * p.out_batch_stride : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty * 4 + i;
        if (r >= p.rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (p.mul) v *= p.mul[z * p.mul_batch_stride + (size_t)r * p.ldm + n];
            if (p.add) v += p.add[(size_t)r * p.N + n];
            if (p.emb) {
                long long lab = p.labels[r];
                if (lab < 0 || lab >= p.num_emb) { if (p.err) atomicOr(p.err, 1); lab = 0; }
                v += p.emb[(size_t)lab * p.N + n];
            }
            out[(size_t)r * p.ldo + n] = v;
            if (out_silu) out_silu[(size_t)r * p.ldo + n] = v / (1.f + expf(-v));
            if (p.out_silu_split) {
                const float sv = v / (1.f + expf(-v)), hi = tf32_round(sv);
                float* d = p.out_silu_split + ((size_t)(r / 64) * 128 + (r % 64)) * p.N + n;
                d[0] = hi; d[(size_t)64 * p.N] = sv - hi;
            }
        }
    }
}

// out = sum over K splits (fixed order) + bias (+ add) (+ embedding row); optional silu(out)
__global__ void small_linear_finalize_kernel(SmallLinear p)
{
    pdl_wait();
    pdl_launch_dependents();
    const int z = blockIdx.z;
    const size_t total = (size_t)p.rows * p.N;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = (int)(i / p.N), n = (int)(i % p.N);
    const float* part = p.partial + (size_t)z * p.ksplit * total + i;
    float v = 0.f;
    for (int s = 0; s < p.ksplit; ++s) v += part[(size_t)s * total];
    if (p.bias) v += p.bias[z * p.bias_batch_stride + n];
    if (p.mul) v *= p.mul[z * p.mul_batch_stride + (size_t)r * p.ldm + n];
    if (p.add) v += p.add[(size_t)r * p.N + n];
    if (p.emb) {
        long long lab = p.labels[r];
        if (lab < 0 || lab >= p.num_emb) { if (p.err) atomicOr(p.err, 1); lab = 0; }
        v += p.emb[(size_t)lab * p.N + n];
    }
    p.out[z * p.out_batch_stride + (size_t)r * p.ldo + n] = v;
    if (p.out_silu) p.out_silu[z * p.out_batch_stride + (size_t)r * p.ldo + n] = v / (1.f + expf(-v));
    if (p.out_silu_split) {
        const float sv = v / (1.f + expf(-v)), hi = tf32_round(sv);
        float* d = p.out_silu_split + ((size_t)(r / 64) * 128 + (r % 64)) * p.N + n;
        d[0] = hi; d[(size_t)64 * p.N] = sv - hi;
    }
}

}  // namespace fitv2
