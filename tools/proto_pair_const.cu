// Prototype (development aid, not part of the library): an FP32 SIMT formulation of the electron-passing pair MLP in
// which ONE THREAD OWNS ONE PAIR and the weights are operands from the constant bank.
//
// The shipped FP32 kernels (tile_gemm, epnn_internal.cuh) give a warp a 32 x 32 output tile with an 8 x 4 register
// tile per thread; both operands come from shared memory, and ncu shows the result: shared-memory wavefronts 58..64 %
// of peak at 44..54 % FMA-pipe activity -- the operand traffic, not the FMA pipe, bounds them.
// Here a thread keeps its pair's 32 pre-activations and 32 accumulators in registers, and every lane of a warp needs the
// SAME weight W[k][c] at the same time, so the weights are passed as a __grid_constant__ kernel parameter (6.4 KB,
// constant bank 0) and appear as immediate constant operands of the FFMAs: acc_c = fma(z_k, c[0][W2[k][c]], acc_c) --
// two register reads per FFMA (no three-register bank conflict), no shared-memory operand traffic, no z stage, no
// shuffles (the w3 dot product is in-thread).  ptxas turns the parameter reads into LDCU.128 (constant bank -> uniform
// registers, one per two packed FFMA2) and emits  FFMA2 R, R.F32, UR.F32x2.HI_LO, R.F32x2.HI_LO  -- the weight PAIR is
// the uniform operand, the activation the broadcast scalar (checked with cuobjdump: 768 FFMA2 + 32 FFMA per direction
// pair, 102 registers, no spills).  Shared memory only holds the staged u | v rows of the window.
// Same test harness as tools/proto_pair_mma.cu: float64 check + pairs/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/proto_pair_const tools/proto_pair_const.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define HID 32
#define EK 16
#define WIN 48                 // atoms of a warp's window (a "bundle")
#define UVS 68                 // row stride of the staged u | v rows: 64 + 4 floats -> rows start in different bank groups
#define TILES_PER_WIN 6
#ifndef NW
#define NW 8
#endif
#ifndef CTAS_PER_SM
#define CTAS_PER_SM 2
#endif

struct PairW { float Cw[EK * HID]; float W2[HID * HID]; float b2[HID]; float w3[HID]; };

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void fma2(u64& d, u64 wpair, float a) { const u64 aa = pack2(a, a); asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(wpair), "l"(aa)); }
struct Args {
    int n_win;
    const float* uv;                        // [n_win*WIN][64]  u | v
    const float* c;                         // [n_tiles*32][EK]
    const unsigned char* li; const unsigned char* lj;
    float* delta;
    int repeat;
};

__global__ void __launch_bounds__(NW * 32, CTAS_PER_SM) pair_const_kernel(const __grid_constant__ PairW W, const Args a) {
    extern __shared__ __align__(16) float s_uv[];             // [NW][WIN * UVS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* uv = s_uv + warp * (WIN * UVS);
    for (int rep = 0; rep < a.repeat; ++rep)
    for (int w = blockIdx.x * NW + warp; w < a.n_win; w += gridDim.x * NW) {
        __syncwarp();
        for (int f = lane; f < WIN * 16; f += 32) {
            const int row = f >> 4, ch4 = f & 15;
            *reinterpret_cast<float4*>(uv + row * UVS + ch4 * 4) = *reinterpret_cast<const float4*>(a.uv + ((size_t)w * WIN + row) * 64 + ch4 * 4);
        }
        __syncwarp();
        for (int tl = 0; tl < TILES_PER_WIN; ++tl) {
            const size_t p = ((size_t)w * TILES_PER_WIN + tl) * 32 + lane;          // lane = pair
            const int ai = a.li[p], aj = a.lj[p];
            // ---- ce = Cw^T c: the pair's 16 coefficients (one 64-byte row) against constant-bank weights
            float cf[EK];
#pragma unroll
            for (int q = 0; q < EK / 4; ++q) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(a.c + p * EK) + q);
                cf[4 * q] = x.x; cf[4 * q + 1] = x.y; cf[4 * q + 2] = x.z; cf[4 * q + 3] = x.w;
            }
            float ce[HID];
            {
                u64 ce2[HID / 2];
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) ce2[o] = 0ull;
#pragma unroll
                for (int k = 0; k < EK; ++k)
#pragma unroll
                    for (int o = 0; o < HID / 2; ++o) fma2(ce2[o], *reinterpret_cast<const u64*>(&W.Cw[k * HID + 2 * o]), cf[k]);
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) unpack2(ce2[o], ce[2 * o], ce[2 * o + 1]);
            }
            float fd = 0.f;
#pragma unroll 1
            for (int dir = 0; dir < 2; ++dir) {
                const float* urow = uv + (dir ? aj : ai) * UVS;            // u of the receiving atom
                const float* vrow = uv + (dir ? ai : aj) * UVS + HID;      // v of the sending atom
                u64 acc2[HID / 2];
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) acc2[o] = pack2(W.b2[2 * o], W.b2[2 * o + 1]);
#pragma unroll
                for (int k4 = 0; k4 < HID / 4; ++k4) {
                    const float4 u4 = *reinterpret_cast<const float4*>(urow + 4 * k4);
                    const float4 v4 = *reinterpret_cast<const float4*>(vrow + 4 * k4);
                    const float z[4] = {fmaxf((ce[4 * k4] + u4.x) + v4.x, 0.f), fmaxf((ce[4 * k4 + 1] + u4.y) + v4.y, 0.f),
                                        fmaxf((ce[4 * k4 + 2] + u4.z) + v4.z, 0.f), fmaxf((ce[4 * k4 + 3] + u4.w) + v4.w, 0.f)};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int o = 0; o < HID / 2; ++o) fma2(acc2[o], *reinterpret_cast<const u64*>(&W.W2[(4 * k4 + kk) * HID + 2 * o]), z[kk]);
                }
                float f = 0.f;
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) {
                    float x, y; unpack2(acc2[o], x, y);
                    f = fmaf(fmaxf(x, 0.f), W.w3[2 * o], f); f = fmaf(fmaxf(y, 0.f), W.w3[2 * o + 1], f);
                }
                fd = dir ? fd - f : f;
            }
            a.delta[p] = 0.5f * fd;
        }
    }
}

static double relu(double x) { return x > 0 ? x : 0; }

int main(int argc, char** argv) {
    const int n_win = argc > 1 ? atoi(argv[1]) : 148 * 8 * 16;
    const size_t n_tiles = (size_t)n_win * TILES_PER_WIN, P = n_tiles * 32, n_at = (size_t)n_win * WIN;
    std::vector<float> uv(n_at * 64), c(P * EK);
    PairW W;
    std::vector<unsigned char> li(P), lj(P);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 65536.f - 0.5f; };
    for (auto& x : uv) x = 1.5f * rnd();
    for (auto& x : c) x = 0.8f * rnd();
    for (auto& x : W.Cw) x = 0.9f * rnd();
    for (auto& x : W.W2) x = 0.7f * rnd();
    for (auto& x : W.b2) x = 0.3f * rnd();
    for (auto& x : W.w3) x = rnd();
    for (size_t p = 0; p < P; ++p) {
        s = s * 1664525u + 1013904223u; li[p] = (unsigned char)((s >> 10) % WIN);
        s = s * 1664525u + 1013904223u; lj[p] = (unsigned char)((s >> 10) % WIN);
    }
    Args a;
    float *d_uv, *d_c, *d_delta; unsigned char *d_li, *d_lj;
    cudaMalloc(&d_uv, uv.size() * 4); cudaMalloc(&d_c, c.size() * 4); cudaMalloc(&d_delta, P * 4); cudaMalloc(&d_li, P); cudaMalloc(&d_lj, P);
    cudaMemcpy(d_uv, uv.data(), uv.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_c, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_li, li.data(), P, cudaMemcpyHostToDevice); cudaMemcpy(d_lj, lj.data(), P, cudaMemcpyHostToDevice);
    a.n_win = n_win; a.uv = d_uv; a.c = d_c; a.li = d_li; a.lj = d_lj; a.delta = d_delta; a.repeat = 1;
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = sizeof(float) * NW * WIN * UVS;
    cudaFuncSetAttribute(pair_const_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = sm * CTAS_PER_SM;
    pair_const_kernel<<<grid, NW * 32, smem>>>(W, a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> delta(P);
    cudaMemcpy(delta.data(), d_delta, P * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    const size_t step = n_tiles / 64 ? n_tiles / 64 : 1;
    for (size_t tl = 0; tl < n_tiles; tl += step) {
        const size_t w = tl / TILES_PER_WIN;
        for (int r = 0; r < 32; ++r) {
            const size_t p = tl * 32 + r;
            double ce[HID];
            for (int o = 0; o < HID; ++o) { double x = 0; for (int k = 0; k < EK; ++k) x += (double)W.Cw[k * HID + o] * c[p * EK + k]; ce[o] = x; }
            double fd[2];
            for (int dir = 0; dir < 2; ++dir) {
                const size_t ia = w * WIN + (dir ? lj[p] : li[p]), ib = w * WIN + (dir ? li[p] : lj[p]);
                double z[HID];
                for (int o = 0; o < HID; ++o) z[o] = relu(ce[o] + uv[ia * 64 + o] + uv[ib * 64 + HID + o]);
                double f = 0;
                for (int o = 0; o < HID; ++o) { double x = W.b2[o]; for (int k = 0; k < HID; ++k) x += z[k] * W.W2[k * HID + o]; f += relu(x) * W.w3[o]; }
                fd[dir] = f;
            }
            const double ref = 0.5 * (fd[0] - fd[1]);
            max_err = fmax(max_err, fabs(ref - delta[p])); max_ref = fmax(max_ref, fabs(ref));
        }
    }
    printf("check: max |delta - ref64| = %.3e  (max |ref| = %.3f)  -> relative %.2e\n", max_err, max_ref, max_err / max_ref);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    a.repeat = 4;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); pair_const_kernel<<<grid, NW * 32, smem>>>(W, a); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
    }
    const double pairs = (double)P * a.repeat;
    printf("timing (%d warps/CTA, %d CTAs/SM): %.3f ms for %.1f M unordered pairs  -> %.2f G pairs/s, %.1f TFLOP/s (5120 FLOP executed per pair)\n",
           NW, CTAS_PER_SM, best, pairs * 1e-6, pairs / (best * 1e-3) * 1e-9, pairs * 5120.0 / (best * 1e-3) * 1e-12);
    e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess || !(max_err / max_ref < 1e-4);
}
