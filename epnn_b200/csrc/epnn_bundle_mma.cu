// OPTIONAL tensor variant of the electron-passing bundle kernel (option "pair_tensor" = 1, precision 32 only).
//
// Same arithmetic as bundle_kernel<float, NW, EPN> (epnn_bundle.cu; reference charge_gn.py:101-116): per unordered
// e != 0 pair p = (i, j) of a bundle of small systems
//     ce    = (B^T C)^T (B^T e_p)                         16 -> 32   (descriptor coefficients in the rank-16 basis)
//     f_ij  = w3 . relu(W2^T relu(ce + u_i + v_j) + b2)   32 -> 32 -> 1,   f_ji with i and j swapped
//     delta = 0.5 (f_ij - f_ji) * is_near_p
// but both matrix products run on the warp-level tensor path (mma.sync.m16n8k8, TF32 inputs, FP32 accumulation) with
// the 3xTF32 error-compensated split  x = hi + lo  (hi = x with the low 13 mantissa bits cleared, lo = x - hi exactly):
// hi*hi + lo*hi + hi*lo.  Everything is fed from registers: the weights live in registers as hi / lo B fragments for
// the whole kernel, the A operand of the first product is loaded from global memory directly in fragment layout (the
// 64-byte coefficient rows are fully used), and the A operand of the second product is what the first product, the u / v
// gathers and the ReLU leave in the C-fragment layout -- inside every block of 8 the k index is permuted (fragment
// position t <-> k = 2t, position t + 4 <-> k = 2t + 1) and W2's rows are loaded with the same permutation, so there is
// no shared-memory z stage.  Only the bundle's u | v rows are staged in (warp-private) shared memory.
// Measured as a stand-alone prototype (tools/proto_pair_mma.cu): 1.57x the pairs/s of the FP32 SIMT kernel, relative
// error 7e-7 against float64.  FP32 SIMT stays the default (BASELINE north_star).
#include "epnn_internal.cuh"

#define MMA_NW 8
#define UVS 72                 // row stride of the staged u | v rows: 64 + 8 floats spreads the rows over the banks

struct EpnMmaArgs {
    int n_bundles; const int2* bundle; int* work_counter;
    const int* ustart; const int* pair_i; const int* pair_j; const unsigned char* near; const float* e;
    const float* u; const float* v;
    const float* Cw; const float* W2; const float* b2; const float* w3;       // [16][32], [32][32], [32], [32]
    float* delta;
};

#ifdef EPNN_CPU_EMU
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) { emu_mma_m16n8k8_tf32(d, a, b); }
#else
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__global__ void __launch_bounds__(MMA_NW * 32, 1) bundle_epn_mma_kernel(const EpnMmaArgs a) {
#ifdef EPNN_CPU_EMU
    float* s_uv = emu_smem;
#else
    extern __shared__ __align__(16) float s_uv[];             // [MMA_NW][BUNDLE_ATOMS * UVS]
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* uv = s_uv + warp * (BUNDLE_ATOMS * UVS);
    // ---- weights -> register-resident B fragments (b0 = (k position t, column g), b1 = (k position t + 4, column g))
    unsigned ch[2][4][2], cl[2][4][2];      // first product, natural k order:   position t <-> k = 8 ks + t
    unsigned wh[4][4][2], wl[4][4][2];      // second product, permuted k order: position t <-> k = 8 ks + 2t, t + 4 <-> 8 ks + 2t + 1
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            split_tf32(a.Cw[(8 * ks + t) * HID + 8 * n + g], ch[ks][n][0], cl[ks][n][0]);
            split_tf32(a.Cw[(8 * ks + t + 4) * HID + 8 * n + g], ch[ks][n][1], cl[ks][n][1]);
        }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            split_tf32(a.W2[(8 * ks + 2 * t) * HID + 8 * n + g], wh[ks][n][0], wl[ks][n][0]);
            split_tf32(a.W2[(8 * ks + 2 * t + 1) * HID + 8 * n + g], wh[ks][n][1], wl[ks][n][1]);
        }
    float2 b2v[4], w3v[4];                  // the thread's 8 output columns: 8n + 2t, 8n + 2t + 1
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        b2v[n] = *reinterpret_cast<const float2*>(a.b2 + 8 * n + 2 * t);
        w3v[n] = *reinterpret_cast<const float2*>(a.w3 + 8 * n + 2 * t);
    }

    auto grab = [&]() {                     // dynamic bundle queue, as in bundle_kernel
        int x = 0;
        if (lane == 0) x = atomicAdd(a.work_counter, 1);
        return __shfl_sync(0xffffffffu, x, 0);
    };
    for (int b = grab(); b < a.n_bundles; b = grab()) {
        const int2 bd = a.bundle[b];
        const int atom0 = bd.x, nat = bd.y;
        const int p0 = a.ustart[atom0], p1 = a.ustart[atom0 + nat];
        __syncwarp();                                              // the previous bundle's rows are no longer read
        for (int f = lane; f < nat * 16; f += 32) {                // stage u | v (coalesced 16-byte loads)
            const int row = f >> 4, c4 = f & 15;
            const float* src = (c4 < 8 ? a.u : a.v) + (int64_t)(atom0 + row) * HID + (c4 & 7) * 4;
            *reinterpret_cast<float4*>(uv + row * UVS + c4 * 4) = *reinterpret_cast<const float4*>(src);
        }
        __syncwarp();
        for (int tb = p0; tb < p1; tb += 32) {
            const int rows = min(32, p1 - tb);
            int my_i = 0, my_j = 0;                                // lane = pair (coalesced); invalid slots point at atom 0
            float my_near = 0.f;
            if (lane < rows) { my_i = a.pair_i[tb + lane] - atom0; my_j = a.pair_j[tb + lane] - atom0; my_near = (float)a.near[tb + lane]; }
            // this thread's four rows: r(m, h) = 16 m + g + 8 h
            int ai[2][2], aj[2][2];
            float nr[2][2];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = 16 * m + g + 8 * h;
                    ai[m][h] = __shfl_sync(0xffffffffu, my_i, r);
                    aj[m][h] = __shfl_sync(0xffffffffu, my_j, r);
                    nr[m][h] = __shfl_sync(0xffffffffu, my_near, r);
                }
            // ---- first product: ce = Cw^T c  (rows beyond the tile re-read its last valid row; their results are dropped)
            float ce[2][4][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int n = 0; n < 4; ++n) { ce[m][n][0] = ce[m][n][1] = ce[m][n][2] = ce[m][n][3] = 0.f; }
                const float* c0 = a.e + (int64_t)(tb + min(16 * m + g, rows - 1)) * EDR + t;
                const float* c1 = a.e + (int64_t)(tb + min(16 * m + g + 8, rows - 1)) * EDR + t;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    unsigned ah[4], al[4];
                    split_tf32(__ldg(c0 + 8 * ks), ah[0], al[0]);          // (row g,     k = t)
                    split_tf32(__ldg(c1 + 8 * ks), ah[1], al[1]);          // (row g + 8, k = t)
                    split_tf32(__ldg(c0 + 8 * ks + 4), ah[2], al[2]);      // (row g,     k = t + 4)
                    split_tf32(__ldg(c1 + 8 * ks + 4), ah[3], al[3]);      // (row g + 8, k = t + 4)
#pragma unroll
                    for (int n = 0; n < 4; ++n) mma_tf32(ce[m][n], ah, ch[ks][n]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) mma_tf32(ce[m][n], al, ch[ks][n]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) mma_tf32(ce[m][n], ah, cl[ks][n]);
                }
            }
            // ---- both directions: dir 0 = i receives from j, dir 1 = j receives from i
            float fd[2][2];                                        // f_ij - f_ji of rows (m, h)
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    float acc[4][4];
#pragma unroll
                    for (int n = 0; n < 4; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
                    const float* urow0 = uv + (dir ? aj[m][0] : ai[m][0]) * UVS;           // u of the receiving atom (row g)
                    const float* vrow0 = uv + (dir ? ai[m][0] : aj[m][0]) * UVS + HID;     // v of the sending atom
                    const float* urow1 = uv + (dir ? aj[m][1] : ai[m][1]) * UVS;           // row g + 8
                    const float* vrow1 = uv + (dir ? ai[m][1] : aj[m][1]) * UVS + HID;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {               // k-step ks consumes the C-layout values of n-tile ks
                        const float2 u0 = *reinterpret_cast<const float2*>(urow0 + 8 * ks + 2 * t);
                        const float2 v0 = *reinterpret_cast<const float2*>(vrow0 + 8 * ks + 2 * t);
                        const float2 u1 = *reinterpret_cast<const float2*>(urow1 + 8 * ks + 2 * t);
                        const float2 v1 = *reinterpret_cast<const float2*>(vrow1 + 8 * ks + 2 * t);
                        unsigned ah[4], al[4];
                        split_tf32(fmaxf((ce[m][ks][0] + u0.x) + v0.x, 0.f), ah[0], al[0]);    // (row g,     k = 2t)
                        split_tf32(fmaxf((ce[m][ks][2] + u1.x) + v1.x, 0.f), ah[1], al[1]);    // (row g + 8, k = 2t)
                        split_tf32(fmaxf((ce[m][ks][1] + u0.y) + v0.y, 0.f), ah[2], al[2]);    // (row g,     k = 2t + 1)
                        split_tf32(fmaxf((ce[m][ks][3] + u1.y) + v1.y, 0.f), ah[3], al[3]);    // (row g + 8, k = 2t + 1)
#pragma unroll
                        for (int n = 0; n < 4; ++n) mma_tf32(acc[n], ah, wh[ks][n]);
#pragma unroll
                        for (int n = 0; n < 4; ++n) mma_tf32(acc[n], al, wh[ks][n]);
#pragma unroll
                        for (int n = 0; n < 4; ++n) mma_tf32(acc[n], ah, wl[ks][n]);
                    }
                    float s0 = 0.f, s1 = 0.f;                      // rows g, g + 8: the thread's 8 columns, fixed order
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        s0 = fmaf(fmaxf(acc[n][0] + b2v[n].x, 0.f), w3v[n].x, s0);
                        s0 = fmaf(fmaxf(acc[n][1] + b2v[n].y, 0.f), w3v[n].y, s0);
                        s1 = fmaf(fmaxf(acc[n][2] + b2v[n].x, 0.f), w3v[n].x, s1);
                        s1 = fmaf(fmaxf(acc[n][3] + b2v[n].y, 0.f), w3v[n].y, s1);
                    }
                    // the four t lanes of a row group hold disjoint columns: butterfly (fixed order -> deterministic)
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                    if (dir == 0) { fd[m][0] = s0; fd[m][1] = s1; } else { fd[m][0] -= s0; fd[m][1] -= s1; }
                }
            }
            if (t == 0) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = 16 * m + g + 8 * h;
                        if (r < rows) a.delta[tb + r] = 0.5f * fd[m][h] * nr[m][h];      // charge_gn.py:116
                    }
            }
        }
    }
}

#ifndef EPNN_CPU_EMU
cudaError_t launch_epn_bundle_mma(const Workspace& w, const StepW<float>& sw, cudaStream_t st, int* nl) {
    if (w.n_bundles == 0) return cudaSuccess;
    EpnMmaArgs ea;
    ea.n_bundles = w.n_bundles; ea.bundle = w.bundle; ea.work_counter = w.work_counter;
    cudaError_t e = cudaMemsetAsync(w.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    ea.ustart = w.ustart; ea.pair_i = w.pair_i; ea.pair_j = w.pair_j; ea.near = w.near; ea.e = w.e;
    ea.u = (const float*)w.u; ea.v = (const float*)w.v;
    ea.Cw = sw.Cw; ea.W2 = sw.W2; ea.b2 = sw.b2; ea.w3 = sw.W3;
    ea.delta = (float*)w.delta;
    const size_t smem = sizeof(float) * MMA_NW * BUNDLE_ATOMS * UVS;
    e = cudaFuncSetAttribute(bundle_epn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = div_up(w.n_bundles, MMA_NW);
    if (grid > w.sm_count) grid = w.sm_count;
    bundle_epn_mma_kernel<<<grid, MMA_NW * 32, smem, st>>>(ea);
    ++*nl;
    return cudaGetLastError();
}
#endif
