// Micro-benchmarks for design decisions (development aid): FP32 FFMA vs packed FFMA2 issue rates,
// mixed FFMA2 + LDS.128 loops.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], a, b);
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}
// 3-register-operand FFMA with distinct multiplicands (GEMM-like: acc += a_s * w_c)
__global__ void __launch_bounds__(256) k_ffma_gemm(float* out, const float* in) {
    float acc[8][4];
    for (int s = 0; s < 8; ++s) for (int c = 0; c < 4; ++c) acc[s][c] = 0.f;
    float a[8], w[4];
    for (int s = 0; s < 8; ++s) a[s] = in[threadIdx.x + s];
    for (int c = 0; c < 4; ++c) w[c] = in[threadIdx.x + 8 + c];
#pragma unroll 4
    for (int i = 0; i < ITERS / 4; ++i) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[s][c] = fmaf(a[s], w[c], acc[s][c]);
        a[i & 7] += 1.0f;
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int c = 0; c < 4; ++c) s += acc[i][c];
    if (s == 123.456f) out[0] = s;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b) {
    unsigned long long x[8];
    for (int i = 0; i < 8; ++i) x[i] = pack(threadIdx.x + i, threadIdx.x - i);
    const unsigned long long A = pack(a, a * 0.5f), B = pack(b, b * 2.f);
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[j]) : "l"(A), "l"(B));
    unsigned long long s = 0; for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x123456789ull) out[0] = 1.f;
}
// GEMM-like FFMA2: acc2[sp][c] += a2[sp] * wdup[c]  (sp = slot pair)
__global__ void __launch_bounds__(256) k_ffma2_gemm(float* out, const float* in) {
    unsigned long long acc[4][4], a[4], w[4];
    for (int s = 0; s < 4; ++s) for (int c = 0; c < 4; ++c) acc[s][c] = 0ull;
    for (int s = 0; s < 4; ++s) a[s] = pack(in[threadIdx.x + s], in[threadIdx.x + s + 1]);
    for (int c = 0; c < 4; ++c) w[c] = pack(in[threadIdx.x + 8 + c], in[threadIdx.x + 8 + c]);
#pragma unroll 4
    for (int i = 0; i < ITERS / 4; ++i) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) fma2(acc[s][c], a[s], w[c]);
        a[i & 3] += 1ull;
    }
    unsigned long long s = 0; for (int i = 0; i < 4; ++i) for (int c = 0; c < 4; ++c) s ^= acc[i][c];
    if (s == 0x123456789ull) out[0] = 1.f;
}
// FFMA2 GEMM inner loop with LDS.128 operands: 8 slot x 4 col tile, W duplicated in smem; measures LDS+FFMA2 co-issue
template <int MODE>
__global__ void __launch_bounds__(256) k_tile(float* out, const float* in, int iters) {
    __shared__ __align__(16) float sA[8 * 32 * 32];      // per warp [32][32]
    __shared__ __align__(16) float sW[32 * 64];          // [k][c][2] duplicated
    for (int t = threadIdx.x; t < 8 * 32 * 32; t += 256) sA[t] = in[t & 1023];
    for (int t = threadIdx.x; t < 32 * 64; t += 256) sW[t] = in[t & 1023];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pg = lane >> 3, og = lane & 7;
    const float* at = sA + warp * 1024 + pg * 8 * 32;
    if (MODE == 0) {          // scalar FFMA, 12 LDS.128 per 128 FFMA (the shipped tile_gemm)
        float acc[8][4];
        for (int s = 0; s < 8; ++s) for (int c = 0; c < 4; ++c) acc[s][c] = 0.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll 2
            for (int kc = 0; kc < 8; ++kc) {
                float4 w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = *reinterpret_cast<const float4*>(sW + (kc * 4 + k) * 64 + og * 4);
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float4 a = *reinterpret_cast<const float4*>(at + s * 32 + ((kc ^ pg) << 2));
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[s][0] = fmaf(av[k], w[k].x, acc[s][0]); acc[s][1] = fmaf(av[k], w[k].y, acc[s][1]);
                        acc[s][2] = fmaf(av[k], w[k].z, acc[s][2]); acc[s][3] = fmaf(av[k], w[k].w, acc[s][3]);
                    }
                }
            }
        }
        float s = 0; for (int i = 0; i < 8; ++i) for (int c = 0; c < 4; ++c) s += acc[i][c];
        if (s == 123.456f) out[0] = s;
    } else {                  // FFMA2 over slot pairs; A tile stored [slotpair][k][2]; W duplicated [k][c][2]
        unsigned long long acc[4][4];
        for (int s = 0; s < 4; ++s) for (int c = 0; c < 4; ++c) acc[s][c] = 0ull;
        for (int it = 0; it < iters; ++it) {
#pragma unroll 2
            for (int kc = 0; kc < 16; ++kc) {         // 2 k per chunk
                ulonglong2 w[2][2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    w[k][0] = *reinterpret_cast<const ulonglong2*>(sW + (kc * 2 + k) * 64 + og * 8);
                    w[k][1] = *reinterpret_cast<const ulonglong2*>(sW + (kc * 2 + k) * 64 + og * 8 + 4);
                }
#pragma unroll
                for (int sp = 0; sp < 4; ++sp) {
                    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(at + sp * 64 + ((kc ^ pg) << 2));
                    fma2(acc[sp][0], a.x, w[0][0].x); fma2(acc[sp][1], a.x, w[0][0].y);
                    fma2(acc[sp][2], a.x, w[0][1].x); fma2(acc[sp][3], a.x, w[0][1].y);
                    fma2(acc[sp][0], a.y, w[1][0].x); fma2(acc[sp][1], a.y, w[1][0].y);
                    fma2(acc[sp][2], a.y, w[1][1].x); fma2(acc[sp][3], a.y, w[1][1].y);
                }
            }
        }
        unsigned long long s = 0; for (int i = 0; i < 4; ++i) for (int c = 0; c < 4; ++c) s ^= acc[i][c];
        if (s == 0x123456789ull) out[0] = 1.f;
    }
}
template <typename F> static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    return best;
}
int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in; cudaMalloc(&out, 1 << 20); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
    const int blocks = sm * 16;
    const double nthr = (double)blocks * 256;
    float ms;
    ms = timeit([&] { k_ffma<<<blocks, 256>>>(out, 0.999f, 0.001f); });
    printf("FFMA  (x=x*a+b, 8 chains)        : %.2f TFLOP/s\n", 2.0 * 8 * ITERS * nthr / ms * 1e-9);
    ms = timeit([&] { k_ffma_gemm<<<blocks, 256>>>(out, in); });
    printf("FFMA  (acc+=a*w, 8x4 reg tile)   : %.2f TFLOP/s\n", 2.0 * 32 * (ITERS / 4) * nthr / ms * 1e-9);
    ms = timeit([&] { k_ffma2<<<blocks, 256>>>(out, 0.999f, 0.001f); });
    printf("FFMA2 (x=x*A+B, 8 chains)        : %.2f TFLOP/s\n", 2.0 * 16 * ITERS * nthr / ms * 1e-9);
    ms = timeit([&] { k_ffma2_gemm<<<blocks, 256>>>(out, in); });
    printf("FFMA2 (acc+=a*w, 4x4 pair tile)  : %.2f TFLOP/s\n", 2.0 * 32 * (ITERS / 4) * nthr / ms * 1e-9);
    for (int occ = 1; occ <= 2; ++occ) {
        const int b2 = sm * occ;
        ms = timeit([&] { k_tile<0><<<b2, 256>>>(out, in, 2000); });
        printf("tile FFMA  + LDS (%d CTA/SM)       : %.2f TFLOP/s\n", occ, 2.0 * 32 * 32 * 2000.0 * b2 * 256 / ms * 1e-9);
        ms = timeit([&] { k_tile<1><<<b2, 256>>>(out, in, 2000); });
        printf("tile FFMA2 + LDS dupW (%d CTA/SM)  : %.2f TFLOP/s\n", occ, 2.0 * 32 * 32 * 2000.0 * b2 * 256 / ms * 1e-9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
