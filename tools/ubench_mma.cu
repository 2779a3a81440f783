// Warp-level tensor-core study for the bundle kernels (development aid, not part of the library):
// can the 32 x 32 x 32 layer-2 product of a 32-pair tile run on the legacy warp-level tensor path (mma.sync m16n8k8,
// TF32 inputs, FP32 accumulation) with the 3xTF32 error-compensated split, fed ENTIRELY from registers?
//   * the A operand (z = relu(ce + u_i + v_j), 32 pairs x 32) is built by each thread directly in the MMA A-fragment
//     layout -- with the k index permuted inside every block of 8 so that the C-fragment columns a thread owns after the
//     previous product (2t, 2t+1) are exactly the A-fragment columns it needs (t, t+4): no shared-memory z stage;
//   * the B operand (W2 hi / lo, rows permuted the same way) stays in registers for the whole kernel: 64 registers.
// Kernels: raw    -- issue rate of independent mma.sync.m16n8k8.tf32 (8 accumulator sets per warp);
//          chain  -- per tile: build z hi/lo (32 values per thread: add, relu, mask, subtract), 96 MMAs
//                    (hi*hi + lo*hi + hi*lo), epilogue relu(. + b2) and a row sum.  Reported as FP32-equivalent
//                    TFLOP/s = tiles * 2*32*32*32 / time, to compare with the FFMA2 tile product (tools/ubench2.cu: 51).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256) raw(float* out, int iters) {
    float d[8][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    for (int j = 0; j < 4; ++j) a[j] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + j);
    b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_tf32(d[i], a, b);
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    if (s == 123.456f) out[0] = s;
}

// One warp = one 32-pair tile per iteration.  Thread (g = lane >> 2, t = lane & 3).
// C / z layout per m-tile m (rows 16m + g, 16m + g + 8) and n-tile n (cols 8n + 2t, 8n + 2t + 1): 4 values.
__global__ void __launch_bounds__(256) chain(float* out, const float* in, int iters) {
    const int lane = threadIdx.x & 31, t = lane & 3;
    unsigned bh[4][4][2], bl[4][4][2];          // [k-step][n-tile][2]: W2 hi / lo fragments (register resident)
    for (int k = 0; k < 4; ++k)
        for (int n = 0; n < 4; ++n)
            for (int r = 0; r < 2; ++r) {
                const float w = in[(k * 16 + n * 4 + r * 2 + t + lane) & 1023];
                const unsigned hi = __float_as_uint(w) & 0xFFFFE000u;
                bh[k][n][r] = hi; bl[k][n][r] = __float_as_uint(w - __uint_as_float(hi));
            }
    float b2[4][2];
    for (int n = 0; n < 4; ++n) { b2[n][0] = in[(n * 8 + 2 * t) & 1023]; b2[n][1] = in[(n * 8 + 2 * t + 1) & 1023]; }
    float ce[2][4][4];                          // first-layer pre-activations of the tile, C layout [m][n][4]
    for (int m = 0; m < 2; ++m) for (int n = 0; n < 4; ++n) for (int j = 0; j < 4; ++j) ce[m][n][j] = in[(m * 64 + n * 16 + j * 4 + lane) & 1023];
    float rowsum[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int it = 0; it < iters; ++it) {
        float acc[2][4][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
        const float shift = rowsum[0][0] * 1e-30f;   // loop-carried dependence: the tiles cannot be hoisted
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // A fragment of k-step k = the thread's own C-layout values of "n-tile" k (k permutation, see header):
                // a0 = (row g, col 2t) a1 = (row g+8, col 2t) a2 = (row g, col 2t+1) a3 = (row g+8, col 2t+1)
                unsigned ah[4], al[4];
                const float zsrc[4] = {ce[m][k][0], ce[m][k][2], ce[m][k][1], ce[m][k][3]};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float z = fmaxf(zsrc[j] + shift, 0.f);
                    const unsigned hi = __float_as_uint(z) & 0xFFFFE000u;
                    ah[j] = hi; al[j] = __float_as_uint(z - __uint_as_float(hi));
                }
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    mma_tf32(acc[m][n], ah, bh[k][n]);
                    mma_tf32(acc[m][n], al, bh[k][n]);
                    mma_tf32(acc[m][n], ah, bl[k][n]);
                }
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                rowsum[m][0] += fmaxf(acc[m][n][0] + b2[n][0], 0.f) + fmaxf(acc[m][n][1] + b2[n][1], 0.f);
                rowsum[m][1] += fmaxf(acc[m][n][2] + b2[n][0], 0.f) + fmaxf(acc[m][n][3] + b2[n][1], 0.f);
            }
        }
    }
    const float s = rowsum[0][0] + rowsum[0][1] + rowsum[1][0] + rowsum[1][1];
    if (s == 123.456f) out[0] = s;
}

template <typename F> static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in; cudaMalloc(&out, 64); cudaMalloc(&in, 4096);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (float)((i * 37) % 101) / 101.f - 0.4f;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 4096;
    for (int bps = 1; bps <= 4; bps *= 2) {          // 8, 16, 32 warps per SM
        const int blocks = sm * bps;
        const float ms = time_ms([&] { raw<<<blocks, 256>>>(out, iters); });
        const double mmas = (double)blocks * 8 * iters * 8;
        printf("raw    %2d warps/SM: %.3f ms  %.1f TFLOP/s TF32 (m16n8k8 = 2048 MAC)  %.2f mma/clk/SM at 1.965 GHz\n", bps * 8, ms,
               mmas * 2048 * 2 / (ms * 1e-3) * 1e-12, mmas / sm / (ms * 1e-3 * 1.965e9));
    }
    for (int bps = 1; bps <= 2; ++bps) {             // 8, 16 warps per SM (the bundle kernels run 8)
        const int blocks = sm * bps;
        const float ms = time_ms([&] { chain<<<blocks, 256>>>(out, in, iters / 4); });
        const double tiles = (double)blocks * 8 * (iters / 4);
        printf("chain  %2d warps/SM: %.3f ms  %.1f TFLOP/s FP32-equivalent (3xTF32, 96 mma per 32x32x32 tile)  %.1f Gtile/s\n", bps * 8, ms,
               tiles * 2.0 * 32 * 32 * 32 / (ms * 1e-3) * 1e-12, tiles / (ms * 1e-3) * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
