/* Plain-C client of the drop-in boundary (include/fitv2_b200.h): no Python, no torch, no C++.
 * Builds the XL/2-width, depth-1 model geometry, binds random weights from host memory, runs one forward over
 * (rows = 4, tokens = 64) twice and checks that the result is finite, non-trivial and bit-identical between the runs.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include tools/cabi_client.c -o /tmp/cabi_client \
 *       -L fitv2_b200 -lfitv2_b200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/fitv2_b200
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "fitv2_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, fitv2_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static uint32_t rng = 12345u;
static float frand(void) { rng = rng * 1664525u + 1013904223u; return ((rng >> 8) / 8388608.0f - 1.0f) * 0.05f; }
static uint16_t bf16(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }

static int upload(fitv2_handle* h, int slot, int64_t n, int is16) {
    void* dev = NULL;
    size_t bytes = (size_t)n * (is16 ? 2 : 4);
    void* host = malloc(bytes);
    for (int64_t i = 0; i < n; ++i) { if (is16) ((uint16_t*)host)[i] = bf16(frand()); else ((float*)host)[i] = frand(); }
    CU(cudaMalloc(&dev, bytes));
    CU(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice));
    free(host);
    CK(fitv2_bind_weight(h, slot, dev, n));
    return 0;
}

int main(void) {
    const int D = 1152, L = 1, H = 16, DH = 72, HM = 3072, LORA = 288, C = 16, NE = 1001, R = 4, N = 64;
    fitv2_config cfg = {D, L, H, DH, HM, LORA, C, NE, FITV2_OPERAND_BF16, 1.0f, 1.0f,
                        /*out_channels*/ 0, FITV2_ADALN_LORA, FITV2_NORM_LAYERNORM, FITV2_NORM_LAYERNORM, FITV2_NORM_LAYERNORM, /*channels_first*/ 0};
    fitv2_handle* h = NULL;
    CK(fitv2_create(&cfg, &h));
    const int64_t numel[FITV2_W_COUNT] = {
        (int64_t)D * C, D, (int64_t)D * 256, D, (int64_t)D * D, D, (int64_t)NE * D, (int64_t)6 * D * D, 6 * D,
        (int64_t)L * LORA * D, L * LORA, (int64_t)L * 6 * D * LORA, L * 6 * D, (int64_t)2 * D * D, 2 * D, (int64_t)C * D, C,
        (int64_t)L * 3 * D * D, L * 3 * D, (int64_t)L * D * D, L * D, (int64_t)L * 2 * HM * D, L * 2 * HM, (int64_t)L * D * HM, L * D,
        DH / 4, DH / 4};
    for (int s = 0; s < FITV2_W_COUNT; ++s) {
        if (numel[s] == 0) continue;                          /* slots of the model variants: unused by this configuration */
        const int is16 = s == FITV2_W_QKV_W || s == FITV2_W_PROJ_W || s == FITV2_W_GATEUP_W || s == FITV2_W_FC2_W;
        if (upload(h, s, numel[s], is16)) return 1;
    }
    const int64_t ws_bytes = fitv2_workspace_bytes(h, R, N);
    if (ws_bytes <= 0) { fprintf(stderr, "workspace query failed: %s\n", fitv2_last_error()); return 1; }
    void* ws = NULL;
    CU(cudaMalloc(&ws, (size_t)ws_bytes));
    CU(cudaMemset(ws, 0xFF, (size_t)ws_bytes));
    CK(fitv2_set_workspace(h, ws, ws_bytes));

    float hx[R * N * 16], ht[R], hmask[R * N], hout[2][R * N * 16];
    int64_t hy[R], hgrid[R * 2 * N];
    for (int i = 0; i < R * N * 16; ++i) hx[i] = frand() * 20.0f;
    for (int r = 0; r < R; ++r) {
        ht[r] = 0.1f + 0.2f * r; hy[r] = r == R - 1 ? 1000 : 7 * r;
        for (int n = 0; n < N; ++n) { hmask[r * N + n] = (r == 1 && n >= 48) ? 0.0f : 1.0f; hgrid[(r * 2 + 0) * N + n] = n % 8; hgrid[(r * 2 + 1) * N + n] = n / 8; }
    }
    float *dx, *dt, *dmask, *dout; int64_t *dy, *dgrid;
    CU(cudaMalloc((void**)&dx, sizeof hx)); CU(cudaMalloc((void**)&dt, sizeof ht)); CU(cudaMalloc((void**)&dmask, sizeof hmask));
    CU(cudaMalloc((void**)&dout, sizeof hout[0])); CU(cudaMalloc((void**)&dy, sizeof hy)); CU(cudaMalloc((void**)&dgrid, sizeof hgrid));
    CU(cudaMemcpy(dx, hx, sizeof hx, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dt, ht, sizeof ht, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dmask, hmask, sizeof hmask, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dy, hy, sizeof hy, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dgrid, hgrid, sizeof hgrid, cudaMemcpyHostToDevice));
    cudaStream_t st;
    CU(cudaStreamCreate(&st));
    for (int run = 0; run < 2; ++run) {
        CK(fitv2_forward(h, dx, R, dt, dy, dgrid, dmask, dout, R, N, st));
        CU(cudaMemcpyAsync(hout[run], dout, sizeof hout[0], cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    double sum = 0.0, amax = 0.0; int bad = 0, pad_nonzero = 0;
    for (int i = 0; i < R * N * 16; ++i) {
        const float v = hout[0][i];
        if (!isfinite(v)) ++bad;
        sum += v; if (fabs(v) > amax) amax = fabs(v);
        if (hmask[i / 16] == 0.0f && v != 0.0f) ++pad_nonzero;
    }
    const int same = memcmp(hout[0], hout[1], sizeof hout[0]) == 0;
    /* the fused CFG + Euler update on the first half */
    float *dz; CU(cudaMalloc((void**)&dz, sizeof(float) * (R / 2) * N * 16));
    CU(cudaMemcpy(dz, hx, sizeof(float) * (R / 2) * N * 16, cudaMemcpyHostToDevice));
    CK(fitv2_cfg_euler(dz, dout, 1.5f, 0.004f, NULL, R / 2, N, 16, st));
    CU(cudaStreamSynchronize(st));
    printf("%s | launches %lld | sum %.6f max|v| %.4f | non-finite %d | pad rows non-zero %d | run-to-run identical %d\n",
           fitv2_version(), (long long)fitv2_kernel_launches(h), sum, amax, bad, pad_nonzero, same);
    fitv2_destroy(h);
    return (bad || pad_nonzero || !same || amax == 0.0) ? 2 : 0;
}
