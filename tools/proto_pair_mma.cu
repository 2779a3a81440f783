// Prototype (development aid, not part of the library): the electron-passing pair MLP of a 32-pair tile on the
// warp-level tensor path, fed from registers -- the candidate for an opt-in tensor variant of the bundle kernels.
//
// Per unordered pair p = (i, j) with descriptor coefficients c_p[16] (epnn_bundle.cu, EPN variant):
//     ce    = Cw^T c_p                                   (16 -> 32)
//     f_ij  = w3 . relu(W2^T relu(ce + u_i + v_j) + b2)  (32 -> 32 -> 1),  f_ji likewise with i, j swapped
//     delta = 0.5 (f_ij - f_ji)
// Both products run as mma.sync.m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, hi = x with the low 13 mantissa bits
// cleared: hi*hi + lo*hi + hi*lo, FP32 accumulation).  The second product takes its A operand straight from the
// registers the first one (plus the u/v gathers and the ReLU) left in the C-fragment layout: inside every block of 8
// the k index is permuted (fragment position t <-> k = 2t, position t+4 <-> k = 2t+1), and the rows of W2 are loaded
// with the same permutation, so no shared-memory z stage exists.  Weights live in registers as hi/lo B fragments.
// The kernel checks itself against a float64 CPU evaluation and reports pairs/s (the FP32 SIMT bundle kernel does
// 6.3 G unordered pairs/s on a B200 in the same layer).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/proto_pair_mma tools/proto_pair_mma.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define HID 32
#define EK 16
#define WIN 48                 // atoms of a warp's window (a "bundle")
#define UVS 72                 // row stride of the staged u|v rows (64 + 8: spreads the rows over the banks)
#define TILES_PER_WIN 6
#define NW 8

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split(float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

struct Args {
    int n_win;                              // windows; window w owns atoms [w*WIN, (w+1)*WIN) and tiles [w*TILES_PER_WIN, ...)
    const float* uv;                        // [n_win*WIN][64]  u | v
    const float* c;                         // [n_tiles*32][EK]
    const unsigned char* li; const unsigned char* lj;   // [n_tiles*32] local atom indices inside the window
    const float* Cw; const float* W2; const float* b2; const float* w3;   // [EK][32], [32][32], [32], [32]
    float* delta;                           // [n_tiles*32]
    int repeat;
};

__global__ void __launch_bounds__(NW * 32, 1) pair_mma_kernel(const Args a) {
    extern __shared__ __align__(16) float s_uv[];             // [NW][WIN * UVS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* uv = s_uv + warp * (WIN * UVS);
    // ---- weights -> register-resident B fragments (hi / lo)
    unsigned ch[2][4][2], cl[2][4][2];      // Cw:  [k-step][n-tile][b0,b1]   natural k order: b0 = (k = 8ks+t), b1 = (k = 8ks+t+4)
    unsigned wh[4][4][2], wl[4][4][2];      // W2:  permuted k order:          b0 = (k = 8ks+2t), b1 = (k = 8ks+2t+1)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            split(a.Cw[(8 * ks + t) * HID + 8 * n + g], ch[ks][n][0], cl[ks][n][0]);
            split(a.Cw[(8 * ks + t + 4) * HID + 8 * n + g], ch[ks][n][1], cl[ks][n][1]);
        }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            split(a.W2[(8 * ks + 2 * t) * HID + 8 * n + g], wh[ks][n][0], wl[ks][n][0]);
            split(a.W2[(8 * ks + 2 * t + 1) * HID + 8 * n + g], wh[ks][n][1], wl[ks][n][1]);
        }
    float2 b2v[4], w3v[4];                  // the thread's 8 output columns: 8n + 2t, 8n + 2t + 1
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        b2v[n] = *reinterpret_cast<const float2*>(a.b2 + 8 * n + 2 * t);
        w3v[n] = *reinterpret_cast<const float2*>(a.w3 + 8 * n + 2 * t);
    }

    for (int rep = 0; rep < a.repeat; ++rep)
    for (int w = blockIdx.x * NW + warp; w < a.n_win; w += gridDim.x * NW) {
        // ---- stage the window's u | v rows (coalesced float4 loads, padded row stride)
        __syncwarp();
        for (int f = lane; f < WIN * 16; f += 32) {
            const int row = f >> 4, ch4 = f & 15;
            const float4 x = *reinterpret_cast<const float4*>(a.uv + ((size_t)w * WIN + row) * 64 + ch4 * 4);
            *reinterpret_cast<float4*>(uv + row * UVS + ch4 * 4) = x;
        }
        __syncwarp();
        for (int tl = 0; tl < TILES_PER_WIN; ++tl) {
            const size_t p0 = ((size_t)w * TILES_PER_WIN + tl) * 32;
            // rows of this thread: r[m][h] = 16 m + g + 8 h
            int ai[2][2], aj[2][2];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) { ai[m][h] = a.li[p0 + 16 * m + g + 8 * h]; aj[m][h] = a.lj[p0 + 16 * m + g + 8 * h]; }
            // ---- product 1: ce = Cw^T c   (A fragments straight from global memory: 64-byte rows, every byte used)
            float ce[2][4][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int n = 0; n < 4; ++n) { ce[m][n][0] = ce[m][n][1] = ce[m][n][2] = ce[m][n][3] = 0.f; }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const float* c0 = a.c + (p0 + 16 * m + g) * EK + 8 * ks + t;
                    unsigned ah[4], al[4];
                    split(__ldg(c0), ah[0], al[0]);                // (row g,     k = t)
                    split(__ldg(c0 + 8 * EK), ah[1], al[1]);       // (row g + 8, k = t)
                    split(__ldg(c0 + 4), ah[2], al[2]);            // (row g,     k = t + 4)
                    split(__ldg(c0 + 8 * EK + 4), ah[3], al[3]);   // (row g + 8, k = t + 4)
                    // three passes over the four independent accumulators: dependent MMAs are four issues apart
#pragma unroll
                    for (int n = 0; n < 4; ++n) mma_tf32(ce[m][n], ah, ch[ks][n]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) mma_tf32(ce[m][n], al, ch[ks][n]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) mma_tf32(ce[m][n], ah, cl[ks][n]);
                }
            }
            // ---- both directions
            float f[2][2][2];                                      // [dir][m][h]
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    float acc[4][4];
#pragma unroll
                    for (int n = 0; n < 4; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
                    const float* urow0 = uv + (dir ? aj[m][0] : ai[m][0]) * UVS;          // u of the receiving atom, row g
                    const float* vrow0 = uv + (dir ? ai[m][0] : aj[m][0]) * UVS + HID;    // v of the sending atom
                    const float* urow1 = uv + (dir ? aj[m][1] : ai[m][1]) * UVS;          // row g + 8
                    const float* vrow1 = uv + (dir ? ai[m][1] : aj[m][1]) * UVS + HID;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {               // k-step ks consumes the C-layout values of n-tile ks
                        const float2 u0 = *reinterpret_cast<const float2*>(urow0 + 8 * ks + 2 * t);
                        const float2 v0 = *reinterpret_cast<const float2*>(vrow0 + 8 * ks + 2 * t);
                        const float2 u1 = *reinterpret_cast<const float2*>(urow1 + 8 * ks + 2 * t);
                        const float2 v1 = *reinterpret_cast<const float2*>(vrow1 + 8 * ks + 2 * t);
                        unsigned ah[4], al[4];
                        split(fmaxf(ce[m][ks][0] + u0.x + v0.x, 0.f), ah[0], al[0]);   // (row g,     k = 2t)
                        split(fmaxf(ce[m][ks][2] + u1.x + v1.x, 0.f), ah[1], al[1]);   // (row g + 8, k = 2t)
                        split(fmaxf(ce[m][ks][1] + u0.y + v0.y, 0.f), ah[2], al[2]);   // (row g,     k = 2t + 1)
                        split(fmaxf(ce[m][ks][3] + u1.y + v1.y, 0.f), ah[3], al[3]);   // (row g + 8, k = 2t + 1)
#pragma unroll
                        for (int n = 0; n < 4; ++n) mma_tf32(acc[n], ah, wh[ks][n]);
#pragma unroll
                        for (int n = 0; n < 4; ++n) mma_tf32(acc[n], al, wh[ks][n]);
#pragma unroll
                        for (int n = 0; n < 4; ++n) mma_tf32(acc[n], ah, wl[ks][n]);
                    }
                    float s0 = 0.f, s1 = 0.f;                      // rows g, g + 8: the thread's 8 columns
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        s0 = fmaf(fmaxf(acc[n][0] + b2v[n].x, 0.f), w3v[n].x, s0);
                        s0 = fmaf(fmaxf(acc[n][1] + b2v[n].y, 0.f), w3v[n].y, s0);
                        s1 = fmaf(fmaxf(acc[n][2] + b2v[n].x, 0.f), w3v[n].x, s1);
                        s1 = fmaf(fmaxf(acc[n][3] + b2v[n].y, 0.f), w3v[n].y, s1);
                    }
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                    f[dir][m][0] = s0; f[dir][m][1] = s1;
                }
            }
            if (t == 0) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) a.delta[p0 + 16 * m + g + 8 * h] = 0.5f * (f[0][m][h] - f[1][m][h]);
            }
        }
    }
}

static double relu(double x) { return x > 0 ? x : 0; }

int main(int argc, char** argv) {
    const int n_win = argc > 1 ? atoi(argv[1]) : 148 * NW * 16;
    const size_t n_tiles = (size_t)n_win * TILES_PER_WIN, P = n_tiles * 32, n_at = (size_t)n_win * WIN;
    std::vector<float> uv(n_at * 64), c(P * EK), Cw(EK * HID), W2(HID * HID), b2(HID), w3(HID);
    std::vector<unsigned char> li(P), lj(P);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xFFFF) / 65536.f - 0.5f; };
    for (auto& x : uv) x = 1.5f * rnd();
    for (auto& x : c) x = 0.8f * rnd();
    for (auto& x : Cw) x = 0.9f * rnd();
    for (auto& x : W2) x = 0.7f * rnd();
    for (auto& x : b2) x = 0.3f * rnd();
    for (auto& x : w3) x = rnd();
    for (size_t p = 0; p < P; ++p) {
        s = s * 1664525u + 1013904223u; li[p] = (unsigned char)((s >> 10) % WIN);
        s = s * 1664525u + 1013904223u; lj[p] = (unsigned char)((s >> 10) % WIN);
    }
    Args a;
    float *d_uv, *d_c, *d_Cw, *d_W2, *d_b2, *d_w3, *d_delta; unsigned char *d_li, *d_lj;
    cudaMalloc(&d_uv, uv.size() * 4); cudaMalloc(&d_c, c.size() * 4); cudaMalloc(&d_Cw, Cw.size() * 4); cudaMalloc(&d_W2, W2.size() * 4);
    cudaMalloc(&d_b2, 128); cudaMalloc(&d_w3, 128); cudaMalloc(&d_delta, P * 4); cudaMalloc(&d_li, P); cudaMalloc(&d_lj, P);
    cudaMemcpy(d_uv, uv.data(), uv.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_c, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_Cw, Cw.data(), Cw.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_W2, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b2, b2.data(), 128, cudaMemcpyHostToDevice); cudaMemcpy(d_w3, w3.data(), 128, cudaMemcpyHostToDevice);
    cudaMemcpy(d_li, li.data(), P, cudaMemcpyHostToDevice); cudaMemcpy(d_lj, lj.data(), P, cudaMemcpyHostToDevice);
    a.n_win = n_win; a.uv = d_uv; a.c = d_c; a.li = d_li; a.lj = d_lj; a.Cw = d_Cw; a.W2 = d_W2; a.b2 = d_b2; a.w3 = d_w3; a.delta = d_delta; a.repeat = 1;
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = sizeof(float) * NW * WIN * UVS;
    cudaFuncSetAttribute(pair_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pair_mma_kernel<<<sm, NW * 32, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> delta(P);
    cudaMemcpy(delta.data(), d_delta, P * 4, cudaMemcpyDeviceToHost);
    // ---- float64 check on a sample of tiles (first, last, some in between)
    double max_err = 0, max_ref = 0;
    const size_t step = n_tiles / 64 ? n_tiles / 64 : 1;
    for (size_t tl = 0; tl < n_tiles; tl += step) {
        const size_t w = tl / TILES_PER_WIN;
        for (int r = 0; r < 32; ++r) {
            const size_t p = tl * 32 + r;
            double ce[HID];
            for (int o = 0; o < HID; ++o) { double x = 0; for (int k = 0; k < EK; ++k) x += (double)Cw[k * HID + o] * c[p * EK + k]; ce[o] = x; }
            double fd[2];
            for (int dir = 0; dir < 2; ++dir) {
                const size_t ia = w * WIN + (dir ? lj[p] : li[p]), ib = w * WIN + (dir ? li[p] : lj[p]);
                double z[HID];
                for (int o = 0; o < HID; ++o) z[o] = relu(ce[o] + uv[ia * 64 + o] + uv[ib * 64 + HID + o]);
                double f = 0;
                for (int o = 0; o < HID; ++o) { double x = b2[o]; for (int k = 0; k < HID; ++k) x += z[k] * W2[k * HID + o]; f += relu(x) * w3[o]; }
                fd[dir] = f;
            }
            const double ref = 0.5 * (fd[0] - fd[1]);
            max_err = fmax(max_err, fabs(ref - delta[p])); max_ref = fmax(max_ref, fabs(ref));
        }
    }
    printf("check: max |delta - ref64| = %.3e  (max |ref| = %.3f)  -> relative %.2e\n", max_err, max_ref, max_err / max_ref);
    // ---- timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    a.repeat = 4;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); pair_mma_kernel<<<sm, NW * 32, smem>>>(a); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
    }
    const double pairs = (double)P * a.repeat;
    printf("timing: %.3f ms for %.1f M unordered pairs  -> %.2f G pairs/s, %.1f TFLOP/s FP32-equivalent (7424 FLOP per pair)\n", best, pairs * 1e-6,
           pairs / (best * 1e-3) * 1e-9, pairs * 7424.0 / (best * 1e-3) * 1e-12);
    e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess || !(max_err / max_ref < 1e-4);
}
