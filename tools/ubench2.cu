// GEMM-core mapping study (development aid): which register-tile / operand-delivery scheme gets closest to the
// FP32 FMA peak for  out[rows][32] += A[rows][32] * W[32][32]  with A and W in shared memory?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench2 tools/ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void fma2s(u64& d, u64 wp, float a) { const u64 aa = pack2(a, a); asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(wp), "l"(aa)); }
__device__ __forceinline__ void fma2p(u64& d, u64 ap, u64 wp) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(ap), "l"(wp)); }

#define K 32
// M0: 8 rows x 4 cols per thread (pg = lane>>3 row group, og = lane&7 col group); warp tile 32 x 32.
__global__ void __launch_bounds__(256) m0(float* out, const float* in, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm; float* sA = sm + K * 32 + (threadIdx.x >> 5) * 32 * K;
    for (int t = threadIdx.x; t < K * 32 + 8 * 32 * K; t += 256) sm[t] = in[t & 1023];
    __syncthreads();
    const int lane = threadIdx.x & 31, pg = lane >> 3, og = lane & 7;
    u64 c[8][2];
    for (int s = 0; s < 8; ++s) c[s][0] = c[s][1] = 0;
    const float* arow = sA + pg * 8 * K;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int kc = 0; kc < K / 4; ++kc) {
            u64 w[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(sW + (kc * 4 + k) * 32 + og * 4); w[k][0] = t.x; w[k][1] = t.y; }
            const int kcs = (kc ^ pg) * 4;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float4 a = *reinterpret_cast<const float4*>(arow + s * K + kcs);
                fma2s(c[s][0], w[0][0], a.x); fma2s(c[s][1], w[0][1], a.x); fma2s(c[s][0], w[1][0], a.y); fma2s(c[s][1], w[1][1], a.y);
                fma2s(c[s][0], w[2][0], a.z); fma2s(c[s][1], w[2][1], a.z); fma2s(c[s][0], w[3][0], a.w); fma2s(c[s][1], w[3][1], a.w);
            }
        }
    }
    u64 x = 0; for (int s = 0; s < 8; ++s) x ^= c[s][0] ^ c[s][1];
    if (x == 0x123456789ull) out[0] = 1.f;
}
// M1: 8 rows x 8 cols per thread (rg = lane>>2: 8 row groups, cg = lane&3: 4 col groups); warp tile 64 x 32.
__global__ void __launch_bounds__(256) m1(float* out, const float* in, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm; float* sA = sm + K * 32 + (threadIdx.x >> 5) * 64 * K;
    for (int t = threadIdx.x; t < K * 32 + 8 * 64 * K; t += 256) sm[t] = in[t & 1023];
    __syncthreads();
    const int lane = threadIdx.x & 31, rg = lane >> 2, cg = lane & 3;
    u64 c[8][4];
    for (int s = 0; s < 8; ++s) for (int q = 0; q < 4; ++q) c[s][q] = 0;
    const float* arow = sA + rg * 8 * K;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int kc = 0; kc < K / 4; ++kc) {
            u64 w[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const ulonglong2 t0 = *reinterpret_cast<const ulonglong2*>(sW + (kc * 4 + k) * 32 + cg * 8);
                const ulonglong2 t1 = *reinterpret_cast<const ulonglong2*>(sW + (kc * 4 + k) * 32 + cg * 8 + 4);
                w[k][0] = t0.x; w[k][1] = t0.y; w[k][2] = t1.x; w[k][3] = t1.y;
            }
            const int kcs = (kc ^ rg) * 4;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float4 a = *reinterpret_cast<const float4*>(arow + s * K + kcs);
                const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int q = 0; q < 4; ++q) fma2s(c[s][q], w[k][q], av[k]);
            }
        }
    }
    u64 x = 0; for (int s = 0; s < 8; ++s) for (int q = 0; q < 4; ++q) x ^= c[s][q];
    if (x == 0x123456789ull) out[0] = 1.f;
}
// M2/M3: row-per-thread: thread owns R rows x all 32 cols; A row in registers; W rows by warp-broadcast LDS.128.
template <int R>
__global__ void __launch_bounds__(256, 1) m23(float* out, const float* in, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm; float* sA = sm + K * 32 + (threadIdx.x >> 5) * 32 * R * (K + 4);
    for (int t = threadIdx.x; t < K * 32 + 8 * 32 * R * (K + 4); t += 256) sm[t] = in[t & 1023];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    u64 c[R][16];
    for (int r = 0; r < R; ++r) for (int q = 0; q < 16; ++q) c[r][q] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int kc = 0; kc < K / 4; ++kc) {
            float a[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 t = *reinterpret_cast<const float4*>(sA + (r * 32 + lane) * (K + 4) + kc * 4);
                a[r][0] = t.x; a[r][1] = t.y; a[r][2] = t.z; a[r][3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(sW + (kc * 4 + k) * 32 + q4 * 4);
#pragma unroll
                    for (int r = 0; r < R; ++r) { fma2s(c[r][q4 * 2], w.x, a[r][k]); fma2s(c[r][q4 * 2 + 1], w.y, a[r][k]); }
                }
            }
        }
    }
    u64 x = 0; for (int r = 0; r < R; ++r) for (int q = 0; q < 16; ++q) x ^= c[r][q];
    if (x == 0x123456789ull) out[0] = 1.f;
}
// M4: 8 x 4 tile with this thread's W columns (32 x 4) held in registers; only A comes from smem.
__global__ void __launch_bounds__(256, 1) m4(float* out, const float* in, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm; float* sA = sm + K * 32 + (threadIdx.x >> 5) * 32 * K;
    for (int t = threadIdx.x; t < K * 32 + 8 * 32 * K; t += 256) sm[t] = in[t & 1023];
    __syncthreads();
    const int lane = threadIdx.x & 31, pg = lane >> 3, og = lane & 7;
    u64 w[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) { const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(sW + k * 32 + og * 4); w[k][0] = t.x; w[k][1] = t.y; }
    u64 c[8][2];
    for (int s = 0; s < 8; ++s) c[s][0] = c[s][1] = 0;
    const float* arow = sA + pg * 8 * K;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kc = 0; kc < K / 4; ++kc) {
            const int kcs = (kc ^ pg) * 4;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float4 a = *reinterpret_cast<const float4*>(arow + s * K + kcs);
                fma2s(c[s][0], w[kc * 4][0], a.x); fma2s(c[s][1], w[kc * 4][1], a.x); fma2s(c[s][0], w[kc * 4 + 1][0], a.y); fma2s(c[s][1], w[kc * 4 + 1][1], a.y);
                fma2s(c[s][0], w[kc * 4 + 2][0], a.z); fma2s(c[s][1], w[kc * 4 + 2][1], a.z); fma2s(c[s][0], w[kc * 4 + 3][0], a.w); fma2s(c[s][1], w[kc * 4 + 3][1], a.w);
            }
        }
    }
    u64 x = 0; for (int s = 0; s < 8; ++s) x ^= c[s][0] ^ c[s][1];
    if (x == 0x123456789ull) out[0] = 1.f;
}
// M5: 16 rows x 4 cols per thread (pg = lane>>3: 4 row groups of 16, og: 8 col groups); warp tile 64 x 32.
__global__ void __launch_bounds__(256) m5(float* out, const float* in, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm; float* sA = sm + K * 32 + (threadIdx.x >> 5) * 64 * K;
    for (int t = threadIdx.x; t < K * 32 + 8 * 64 * K; t += 256) sm[t] = in[t & 1023];
    __syncthreads();
    const int lane = threadIdx.x & 31, pg = lane >> 3, og = lane & 7;
    u64 c[16][2];
    for (int s = 0; s < 16; ++s) c[s][0] = c[s][1] = 0;
    const float* arow = sA + pg * 16 * K;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int kc = 0; kc < K / 4; ++kc) {
            u64 w[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(sW + (kc * 4 + k) * 32 + og * 4); w[k][0] = t.x; w[k][1] = t.y; }
            const int kcs = (kc ^ pg) * 4;
#pragma unroll
            for (int s = 0; s < 16; ++s) {
                const float4 a = *reinterpret_cast<const float4*>(arow + s * K + kcs);
                fma2s(c[s][0], w[0][0], a.x); fma2s(c[s][1], w[0][1], a.x); fma2s(c[s][0], w[1][0], a.y); fma2s(c[s][1], w[1][1], a.y);
                fma2s(c[s][0], w[2][0], a.z); fma2s(c[s][1], w[2][1], a.z); fma2s(c[s][0], w[3][0], a.w); fma2s(c[s][1], w[3][1], a.w);
            }
        }
    }
    u64 x = 0; for (int s = 0; s < 16; ++s) x ^= c[s][0] ^ c[s][1];
    if (x == 0x123456789ull) out[0] = 1.f;
}
template <typename F> static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    return best;
}
int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in; cudaMalloc(&out, 1 << 20); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
    const int iters = 2000;
    auto report = [&](const char* name, double rows_per_warp, float ms, int ctas) {
        printf("%-46s %d CTA/SM: %6.2f TFLOP/s\n", name, ctas, 2.0 * rows_per_warp * K * 32 * iters * 8.0 * sm * ctas / ms * 1e-9);
    };
#define RUN(kern, name, rows, smem)                                                              \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);               \
    for (int ctas = 1; ctas <= 2; ++ctas) {                                                      \
        if ((size_t)(smem) * ctas > 220 * 1024) break;                                           \
        float ms = timeit([&] { kern<<<sm * ctas, 256, smem>>>(out, in, iters); });              \
        report(name, rows, ms, ctas);                                                            \
    }
    RUN(m0, "M0 8x4 tile, W+A smem (shipped)", 32, (K * 32 + 8 * 32 * K) * 4)
    RUN(m1, "M1 8x8 tile", 64, (K * 32 + 8 * 64 * K) * 4)
    RUN(m23<1>, "M2 row-per-thread, A regs, W bcast", 32, (K * 32 + 8 * 32 * (K + 4)) * 4)
    RUN(m23<2>, "M3 2 rows per thread, A regs, W bcast", 64, (K * 32 + 8 * 64 * (K + 4)) * 4)
    RUN(m4, "M4 8x4 tile, W in regs", 32, (K * 32 + 8 * 32 * K) * 4)
    RUN(m5, "M5 16x4 tile", 64, (K * 32 + 8 * 64 * K) * 4)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
