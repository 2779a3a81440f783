// Weight-delivery study for the pair-per-thread kernels (development aid): one thread owns one pair slot and multiplies
// its 32 activations by the SAME 32 x 32 weight block as every other thread.  How should the weights reach the FMA pipe?
//   ldcu<B> : __grid_constant__ parameter -> ptxas emits LDCU.128 (uniform registers) + FFMA2 R, R.F32, UR.F32x2, R.F32x2.
//             Inside a loop ptxas rotates only TWO uniform quads, so every pair of FFMA2 waits for its own LDCU.
//             B = slots per thread (register blocking: one quad then feeds 2 B FFMA2).
//   ldc<B>  : same parameter, per-lane (runtime-zero) index -> LDC.64 into ordinary registers, pipelined by ptxas
//   lds<B>  : weights in shared memory, broadcast LDS.128 into ordinary registers
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_uniform tools/ubench_uniform.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void fma2s(u64& d, u64 wp, float a) { const u64 aa = pack2(a, a); asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(wp), "l"(aa)); }

struct Wt { float W2[32 * 32]; float b2[32]; };

template <int B> __device__ __forceinline__ void load_z(float (&z)[B][32], const float* zs, int it, float bias) {
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(zs + ((it + b) & 7) * 36 + 4 * c);
            z[b][4 * c] = fmaxf(x.x + bias, 0.f); z[b][4 * c + 1] = fmaxf(x.y + bias, 0.f); z[b][4 * c + 2] = fmaxf(x.z + bias, 0.f); z[b][4 * c + 3] = fmaxf(x.w + bias, 0.f);
        }
}
template <int B> __device__ __forceinline__ float finish(u64 (&acc)[B][16]) {
    float s = 0.f;
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
        for (int o = 0; o < 16; ++o) { float x, y; unpack2(acc[b][o], x, y); s += fmaxf(x, 0.f) + fmaxf(y, 0.f); }
    return s;
}

template <int NT, int B>
__global__ void __launch_bounds__(NT, 1) k_ldcu(const __grid_constant__ Wt W, const float* in, float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* zs = sm + threadIdx.x * 36 * 0 + (threadIdx.x & 31) * 36;      // 8 rows x 36 per lane slot (shared by warps: read only)
    for (int t = threadIdx.x; t < 32 * 36 + 8 * 36; t += NT) sm[t] = in[t & 1023];
    __syncthreads();
    float s = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; it += B) {
        float z[B][32];
        load_z<B>(z, zs, it, s * 1e-20f);
        u64 acc[B][16];
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[b][o] = pack2(W.b2[2 * o], W.b2[2 * o + 1]);
#pragma unroll
        for (int k = 0; k < 32; ++k)
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const u64 w = *reinterpret_cast<const u64*>(&W.W2[k * 32 + 2 * o]);
#pragma unroll
                for (int b = 0; b < B; ++b) fma2s(acc[b][o], w, z[b][k]);
            }
        s += finish<B>(acc);
    }
    if (s == 123.456f) out[0] = s;
}

//   idx<B>  : like ldcu, but the weight index carries a loop-variant zero.  Without it ptxas hoists as many weight quads as
//             fit the uniform register file out of the loop (15 of 256) and rotates the remaining 241 through ONE or TWO quads.
template <int NT, int B>
__global__ void __launch_bounds__(NT, 1) k_idx(const __grid_constant__ Wt W, const float* in, float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* zs = sm + threadIdx.x * 36 * 0 + (threadIdx.x & 31) * 36;      // 8 rows x 36 per lane slot (shared by warps: read only)
    for (int t = threadIdx.x; t < 32 * 36 + 8 * 36; t += NT) sm[t] = in[t & 1023];
    __syncthreads();
    float s = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; it += B) {
        const int uz = (it >> 28) * 4;      // always 0, but loop-variant: ptxas cannot hoist the weight loads out of the loop
        float z[B][32];
        load_z<B>(z, zs, it, s * 1e-20f);
        u64 acc[B][16];
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[b][o] = pack2(W.b2[2 * o], W.b2[2 * o + 1]);
#pragma unroll
        for (int k = 0; k < 32; ++k)
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const u64 w = *reinterpret_cast<const u64*>(&W.W2[k * 32 + 2 * o + uz]);
#pragma unroll
                for (int b = 0; b < B; ++b) fma2s(acc[b][o], w, z[b][k]);
            }
        s += finish<B>(acc);
    }
    if (s == 123.456f) out[0] = s;
}

template <int NT, int B>
__global__ void __launch_bounds__(NT, 1) k_ldc(const __grid_constant__ Wt W, const float* in, float* out, int iters, int zero) {
    extern __shared__ __align__(16) float sm[];
    float* zs = sm + (threadIdx.x & 31) * 36;
    for (int t = threadIdx.x; t < 32 * 36 + 8 * 36; t += NT) sm[t] = in[t & 1023];
    __syncthreads();
    const int zoff = 2 * zero * (threadIdx.x & 31);
    float s = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; it += B) {
        float z[B][32];
        load_z<B>(z, zs, it, s * 1e-20f);
        u64 acc[B][16];
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[b][o] = pack2(W.b2[2 * o], W.b2[2 * o + 1]);
#pragma unroll
        for (int k = 0; k < 32; ++k)
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const u64 w = *reinterpret_cast<const u64*>(&W.W2[k * 32 + 2 * o + zoff]);
#pragma unroll
                for (int b = 0; b < B; ++b) fma2s(acc[b][o], w, z[b][k]);
            }
        s += finish<B>(acc);
    }
    if (s == 123.456f) out[0] = s;
}

template <int NT, int B>
__global__ void __launch_bounds__(NT, 1) k_lds(const Wt* Wg, const float* in, float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* zs = sm + (threadIdx.x & 31) * 36;
    float* sW = sm + 40 * 36;
    for (int t = threadIdx.x; t < 32 * 36 + 8 * 36; t += NT) sm[t] = in[t & 1023];
    for (int t = threadIdx.x; t < 1024 + 32; t += NT) sW[t] = t < 1024 ? Wg->W2[t] : Wg->b2[t - 1024];
    __syncthreads();
    float s = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; it += B) {
        float z[B][32];
        load_z<B>(z, zs, it, s * 1e-20f);
        u64 acc[B][16];
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[b][o] = pack2(sW[1024 + 2 * o], sW[1024 + 2 * o + 1]);
#pragma unroll
        for (int k = 0; k < 32; ++k)
#pragma unroll
            for (int o = 0; o < 16; o += 2) {
                const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(sW + k * 32 + 2 * o);
#pragma unroll
                for (int b = 0; b < B; ++b) { fma2s(acc[b][o], w.x, z[b][k]); fma2s(acc[b][o + 1], w.y, z[b][k]); }
            }
        s += finish<B>(acc);
    }
    if (s == 123.456f) out[0] = s;
}


// Constant-cache capacity: the same product with KD weight rows (KD * 128 bytes walked cyclically once per iteration).
template <int KD> struct Wk { float W2[KD * 32]; float b2[32]; };
template <int NT, int KD>
__global__ void __launch_bounds__(NT, 1) k_cap(const __grid_constant__ Wk<KD> W, const float* in, float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* zs = sm + (threadIdx.x & 31) * 36;
    for (int t = threadIdx.x; t < 32 * 36 + 8 * 36; t += NT) sm[t] = in[t & 1023];
    __syncthreads();
    float s = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float z[1][32];
        load_z<1>(z, zs, it, s * 1e-20f);
        u64 acc[1][16];
#pragma unroll
        for (int o = 0; o < 16; ++o) acc[0][o] = pack2(W.b2[2 * o], W.b2[2 * o + 1]);
#pragma unroll
        for (int k = 0; k < KD; ++k)
#pragma unroll
            for (int o = 0; o < 16; ++o) fma2s(acc[0][o], *reinterpret_cast<const u64*>(&W.W2[k * 32 + 2 * o]), z[0][k & 31]);
        s += finish<1>(acc);
    }
    if (s == 123.456f) out[0] = s;
}
template <int KD> void run_cap(const float* in, float* out, int iters, cudaEvent_t e0, cudaEvent_t e1) {
    static Wk<KD> hW;
    for (int i = 0; i < KD * 32; ++i) hW.W2[i] = 0.001f * (i % 37) - 0.01f;
    for (int i = 0; i < 32; ++i) hW.b2[i] = 0.01f * i;
    const int grid = 148 * 4, NT = 512;
    const size_t smem = sizeof(float) * (40 * 36 + 1024 + 64);
    k_cap<NT, KD><<<grid, NT, smem>>>(hW, in, out, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_cap<NT, KD><<<grid, NT, smem>>>(hW, in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cap: %5d bytes of weights, 16 warps/SM  %8.3f ms  %6.2f TFLOP/s %s\n", KD * 128, ms,
           2.0 * KD * 32 * (double)iters * grid * NT / (ms * 1e-3) * 1e-12, cudaGetErrorString(cudaGetLastError()));
}

#define TIME(name, launch)                                                                                      \
    do {                                                                                                        \
        launch; cudaDeviceSynchronize();                                                                        \
        cudaEventRecord(e0); launch; cudaEventRecord(e1); cudaEventSynchronize(e1);                             \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                                            \
        cudaError_t err = cudaGetLastError();                                                                   \
        printf("%-28s %8.3f ms  %6.2f TFLOP/s %s\n", name, ms, 2.0 * 1024 * (double)iters * grid * nt / (ms * 1e-3) * 1e-12, \
               err == cudaSuccess ? "" : cudaGetErrorString(err));                                              \
    } while (0)

template <int NT> void run_nt(const Wt& hW, const Wt* dW, const float* in, float* out, int iters, cudaEvent_t e0, cudaEvent_t e1) {
    const int grid = 148 * 4, nt = NT;
    const size_t smem = sizeof(float) * (40 * 36 + 1024 + 64);
    char nm[64];
    snprintf(nm, sizeof nm, "ldcu<1> %d warps/SM", NT / 32); TIME(nm, (k_ldcu<NT, 1><<<grid, NT, smem>>>(hW, in, out, iters)));
    snprintf(nm, sizeof nm, "ldcu<2> %d warps/SM", NT / 32); TIME(nm, (k_ldcu<NT, 2><<<grid, NT, smem>>>(hW, in, out, iters)));
    snprintf(nm, sizeof nm, "idx<1>  %d warps/SM", NT / 32); TIME(nm, (k_idx<NT, 1><<<grid, NT, smem>>>(hW, in, out, iters)));
    snprintf(nm, sizeof nm, "idx<2>  %d warps/SM", NT / 32); TIME(nm, (k_idx<NT, 2><<<grid, NT, smem>>>(hW, in, out, iters)));
    if (getenv("SKIP_SLOW")) return;
    snprintf(nm, sizeof nm, "ldc<1>  %d warps/SM", NT / 32); TIME(nm, (k_ldc<NT, 1><<<grid, NT, smem>>>(hW, in, out, iters, 0)));
    snprintf(nm, sizeof nm, "ldc<2>  %d warps/SM", NT / 32); TIME(nm, (k_ldc<NT, 2><<<grid, NT, smem>>>(hW, in, out, iters, 0)));
    snprintf(nm, sizeof nm, "lds<1>  %d warps/SM", NT / 32); TIME(nm, (k_lds<NT, 1><<<grid, NT, smem>>>(dW, in, out, iters)));
    snprintf(nm, sizeof nm, "lds<2>  %d warps/SM", NT / 32); TIME(nm, (k_lds<NT, 2><<<grid, NT, smem>>>(dW, in, out, iters)));
}

int main() {
    Wt hW;
    for (int i = 0; i < 1024; ++i) hW.W2[i] = 0.001f * (i % 37) - 0.01f;
    for (int i = 0; i < 32; ++i) hW.b2[i] = 0.01f * i;
    Wt* dW; cudaMalloc(&dW, sizeof(Wt)); cudaMemcpy(dW, &hW, sizeof(Wt), cudaMemcpyHostToDevice);
    float *in, *out; cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 64);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 0.01f * (i % 91) - 0.3f;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 512;
    run_cap<16>(in, out, iters, e0, e1); run_cap<24>(in, out, iters, e0, e1); run_cap<32>(in, out, iters, e0, e1); run_cap<40>(in, out, iters, e0, e1);
    run_cap<48>(in, out, iters, e0, e1); run_cap<56>(in, out, iters, e0, e1); run_cap<64>(in, out, iters, e0, e1); run_cap<96>(in, out, iters, e0, e1); run_cap<128>(in, out, iters, e0, e1);
    if (getenv("CAP_ONLY")) return 0;
    run_nt<128>(hW, dW, in, out, iters, e0, e1);
    run_nt<256>(hW, dW, in, out, iters, e0, e1);
    run_nt<384>(hW, dW, in, out, iters, e0, e1);
    run_nt<512>(hW, dW, in, out, iters, e0, e1);
    run_nt<768>(hW, dW, in, out, iters, e0, e1);
    return 0;
}
