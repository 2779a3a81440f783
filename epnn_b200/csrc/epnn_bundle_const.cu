// EXPERIMENTAL FP32 variant of both bundle kernels (option "pair_const" = 1, precision 32 only, default OFF).
// Written at the end of round 1 after the GPU budget was spent: it compiles for sm_100a, its inner loops are those of the
// measured prototype tools/proto_pair_const.cu (9.9 G pairs/s, 73 % of the FFMA peak, 1.9e-7 relative to float64), but the
// kernel as a whole has NOT run on a GPU yet -- its tests (tests/test_gpu_pair_const.py) are gated behind
// EPNN_TEST_EXPERIMENTAL=1 until it has.
//
// Same arithmetic and the same per-bundle flow as bundle_kernel<float, NW, EPN> (epnn_bundle.cu; reference
// charge_gn.py:62-70 and :101-116), with a different mapping of the work onto the warp:
//   * ONE THREAD OWNS ONE PAIR SLOT of a 32-slot tile and keeps the slot's 32 first-layer pre-activations and its 32
//     second-layer accumulators in registers;
//   * every lane needs the same weight at the same time, so the weights are a __grid_constant__ kernel parameter
//     (6.4 KB, constant bank 0): ptxas loads them into uniform registers (LDCU.128) and emits
//     FFMA2 R, R.F32, UR.F32x2, R.F32x2 -- the weight pair is the uniform operand, the activation the broadcast scalar.
//     No shared-memory operand traffic (the bound of tile_gemm), no z stage, no shuffles in the
//     products;
//   * electron passing: the w3 dot product is in-thread, delta is written by the slot's own lane;
//   * message passing: the 32 message columns of a slot go through a 32 x 33 shared-memory transpose, then lane c adds
//     column c into the bundle's S rows slot by slot, in slot order (runs of equal targets are summed in a register
//     first) -- fixed order, no atomics, bitwise reproducible; scatter_sorted / perm_j are not needed here.
// Shared memory per warp: the bundle's u | v rows (row stride 68 floats, + one row holding b1 for the pad pseudo-atom),
// and for the GNN variant the S accumulators, the transpose tile and the pad weights.
#ifdef EPNN_CPU_EMU
#include "../../tools/emu/cuda_emu.h"      // CPU warp emulation (tests/test_emu_pair_const.py): same kernel body, lanes = host threads
#else
#include "epnn_internal.cuh"
#endif

#ifndef CONST_NW
#define CONST_NW 8
#endif
#ifndef CONST_CTAS_EPN
#define CONST_CTAS_EPN 2
#endif
#define CUVS 68                         // row stride of the staged u | v rows (64 + 4: rows start in different bank groups)
#define PAD_ROW BUNDLE_ATOMS            // extra row: v = b1 (a_j = 0, e = 0), the pad pseudo-atom of the GNN's far tiles

// The only constants the tile loop walks: W2, exactly 4 KB = the level-0 constant cache (tools/ubench_uniform.cu: the
// uniform-operand path drops from 64.7 to 41 TFLOP/s as soon as a loop's constants exceed it).  C rows, b2 and x32 (b1 for
// the GNN variant, w3 for the EPN variant) are staged in shared memory, the kernel arguments come through global memory
// (as parameters ptxas re-reads them from the constant bank inside the loop).
struct alignas(16) PairW { float W2[HID * HID]; };      // read with 8- / 16-byte loads

struct ConstArgs {
    int n_bundles; const int2* bundle; int* work_counter;
    const int* ustart; const int* pair_i; const int* pair_j; const unsigned char* near; const float* e;
    const int* far_off; const unsigned short* far_list;
    const int* far0_off; const unsigned short* far0_list; const unsigned char* far0_w; const int* rep; int dedup;
    const int* atom_sys; const int* sys_off; const int* npad;
    const float* u; const float* v;
    const float* Cw; const float* b2; const float* x32;      // device pointers into the packed weights
    float* S; float* delta;
};

typedef unsigned long long f2_t;
#ifdef EPNN_CPU_EMU      // inline PTX replaced by its definition: two IEEE fma.rn on the packed halves
__device__ __forceinline__ f2_t cpack2(float lo, float hi) { unsigned a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return (f2_t)a | ((f2_t)b << 32); }
__device__ __forceinline__ void cunpack2(f2_t v, float& lo, float& hi) { const unsigned a = (unsigned)v, b = (unsigned)(v >> 32); memcpy(&lo, &a, 4); memcpy(&hi, &b, 4); }
__device__ __forceinline__ void cfma2(f2_t& d, f2_t wpair, float a) {
    float d0, d1, w0, w1;
    cunpack2(d, d0, d1); cunpack2(wpair, w0, w1);
    d = cpack2(fmaf(w0, a, d0), fmaf(w1, a, d1));
}
#else
__device__ __forceinline__ f2_t cpack2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void cunpack2(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void cfma2(f2_t& d, f2_t wpair, float a) {
    const f2_t aa = cpack2(a, a);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(wpair), "l"(aa));
}
#endif

#ifdef EPNN_CPU_EMU
__device__ __forceinline__ void cprefetch_l2(const void*) {}
#else
__device__ __forceinline__ void cprefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// acc[c] = b2[c] + sum_k relu(ce[k] + urow[k] + vrow[k]) * W2[k][c]   (ce == nullptr-like: pass zeros for far slots)
template <bool WITH_CE>
__device__ __forceinline__ void second_layer(const PairW& W, const float* __restrict__ sb2, const float (&ce)[HID], const float* urow, const float* vrow, f2_t (&acc2)[HID / 2]) {
#pragma unroll
    for (int o = 0; o < HID / 2; o += 2) {
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(sb2 + 2 * o);
        acc2[o] = b.x; acc2[o + 1] = b.y;
    }
#pragma unroll
    for (int k4 = 0; k4 < HID / 4; ++k4) {
        const float4 u4 = *reinterpret_cast<const float4*>(urow + 4 * k4);
        const float4 v4 = *reinterpret_cast<const float4*>(vrow + 4 * k4);
        float z[4];
        if (WITH_CE) {
            z[0] = fmaxf((ce[4 * k4] + u4.x) + v4.x, 0.f); z[1] = fmaxf((ce[4 * k4 + 1] + u4.y) + v4.y, 0.f);
            z[2] = fmaxf((ce[4 * k4 + 2] + u4.z) + v4.z, 0.f); z[3] = fmaxf((ce[4 * k4 + 3] + u4.w) + v4.w, 0.f);
        } else {
            z[0] = fmaxf(u4.x + v4.x, 0.f); z[1] = fmaxf(u4.y + v4.y, 0.f); z[2] = fmaxf(u4.z + v4.z, 0.f); z[3] = fmaxf(u4.w + v4.w, 0.f);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int o = 0; o < HID / 2; ++o) cfma2(acc2[o], *reinterpret_cast<const f2_t*>(&W.W2[(4 * k4 + kk) * HID + 2 * o]), z[kk]);
    }
}

// Adds  wgt * relu(acc)  of every slot into S[tgt] (tgt < 0: slot unused).  Lane = slot on entry; lane = column while adding.
__device__ __forceinline__ void add_messages(const f2_t (&acc2)[HID / 2], float wgt, int tgt, float* __restrict__ Mt, float* __restrict__ S, int lane) {
#pragma unroll
    for (int o = 0; o < HID / 2; ++o) {
        float x, y;
        cunpack2(acc2[o], x, y);
        Mt[lane * 33 + 2 * o] = fmaxf(x, 0.f) * wgt;               // bank (lane + column) mod 32: conflict-free
        Mt[lane * 33 + 2 * o + 1] = fmaxf(y, 0.f) * wgt;
    }
    __syncwarp();
    int cur = -1;
    float run = 0.f;
#pragma unroll 4
    for (int s = 0; s < 32; ++s) {                                  // warp-uniform control flow: every lane sees the same targets
        const int t = __shfl_sync(0xffffffffu, tgt, s);
        if (t != cur) {
            if (cur >= 0) S[cur * HID + lane] += run;
            run = 0.f; cur = t;
        }
        if (t >= 0) run += Mt[s * 33 + lane];
    }
    if (cur >= 0) S[cur * HID + lane] += run;
    __syncwarp();
}

template <bool EPN> struct ConstSmem {
    static constexpr int PW = (BUNDLE_ATOMS + 1) * CUVS + (EPN ? 0 : BUNDLE_ATOMS * HID + 32 * 33 + BUNDLE_ATOMS);   // floats per warp
    static constexpr int SHARED = EDR * HID + 2 * HID;                    // C rows, b2, x32 (per CTA)
    static size_t bytes() { return sizeof(float) * ((size_t)CONST_NW * PW + SHARED); }
};

template <bool EPN>
__global__ void __launch_bounds__(CONST_NW * 32, EPN ? CONST_CTAS_EPN : 1) bundle_const_kernel(const __grid_constant__ PairW W, const ConstArgs* __restrict__ ap) {
    const ConstArgs a = *ap;
#ifdef EPNN_CPU_EMU
    float* csm = emu_smem;
#else
    extern __shared__ __align__(16) float csm[];
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sCw = csm;                                              // [EDR][32]
    float* sb2 = csm + EDR * HID;                                  // [32]
    float* sx = sb2 + HID;                                         // [32]  b1 (GNN) / w3 (EPN)
    for (int t = threadIdx.x; t < EDR * HID; t += blockDim.x) sCw[t] = a.Cw[t];
    if (threadIdx.x < HID) { sb2[threadIdx.x] = a.b2[threadIdx.x]; sx[threadIdx.x] = a.x32[threadIdx.x]; }
    __syncthreads();
    float* uv = csm + ConstSmem<EPN>::SHARED + warp * ConstSmem<EPN>::PW;   // [BUNDLE_ATOMS + 1][CUVS]
    float* S = uv + (BUNDLE_ATOMS + 1) * CUVS;                     // [BUNDLE_ATOMS][32]   (GNN)
    float* Mt = S + BUNDLE_ATOMS * HID;                            // [32][33]             (GNN)
    float* padw = Mt + 32 * 33;                                    // [BUNDLE_ATOMS]       (GNN)
    if (!EPN) uv[PAD_ROW * CUVS + HID + lane] = sx[lane];       // v of the pad pseudo-atom = b1 (written once per warp)

    auto grab = [&]() {
        int x = 0;
        if (lane == 0) x = atomicAdd(a.work_counter, 1);
        return __shfl_sync(0xffffffffu, x, 0);
    };
    // every warp already holds the index of its NEXT bundle, to pull that bundle's u / v rows into L2 while it works
    int b = grab();
    int b_next = grab();
    for (; b < a.n_bundles; b = b_next, b_next = grab()) {
        const int2 bd = a.bundle[b];
        const int atom0 = bd.x, nat = bd.y;
        const int p0 = a.ustart[atom0], p1 = a.ustart[atom0 + nat];
        if (b_next < a.n_bundles) {
            const int2 nd = a.bundle[b_next];
            const int nbytes = nd.y * HID * (int)sizeof(float);
            for (int o = lane * 128; o < nbytes; o += 32 * 128) {
                cprefetch_l2(reinterpret_cast<const char*>(a.u + (int64_t)nd.x * HID) + o);
                cprefetch_l2(reinterpret_cast<const char*>(a.v + (int64_t)nd.x * HID) + o);
            }
        }
        __syncwarp();
        // stage u | v (coalesced 16-byte loads), eight loads in flight per lane: with one load per loop trip a bundle waited out
        // 24 L2 latencies one after the other (ncu source view, call77: 17 % of the kernel's stall samples on that one line)
        for (int f0 = 0; f0 < nat * 16; f0 += 8 * 32) {
            float4 x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + q * 32 + lane;
                const int fc = f < nat * 16 ? f : nat * 16 - 1;
                const int row = fc >> 4, c4 = fc & 15;
                x[q] = __ldg(reinterpret_cast<const float4*>((c4 < 8 ? a.u : a.v) + (int64_t)(atom0 + row) * HID + (c4 & 7) * 4));
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = f0 + q * 32 + lane;
                if (f < nat * 16) *reinterpret_cast<float4*>(uv + (f >> 4) * CUVS + (f & 15) * 4) = x[q];
            }
        }
        if (!EPN) {
            for (int f = lane; f < nat * HID; f += 32) S[f] = 0.f;
            for (int r = lane; r < nat; r += 32) {
                const int sys = a.atom_sys[atom0 + r];
                padw[r] = (float)(a.npad[sys] - (a.sys_off[sys + 1] - a.sys_off[sys]));
            }
        }
        __syncwarp();

        // ---------------------------------------------------------------- near tiles: lane = unordered e != 0 pair
        // software prefetch: the next tile's indices and coefficient row are requested as soon as the current tile's
        // first product has consumed its coefficients (same registers), and arrive during the two second-layer products
        int n_i = atom0, n_j = atom0;
        float n_near = 0.f;
        float cf[EDR];
        auto fetch_tile = [&](int t0) {
            n_i = atom0; n_j = atom0; n_near = 0.f;
#pragma unroll
            for (int k = 0; k < EDR; ++k) cf[k] = 0.f;
            if (t0 + lane < p1) {
                n_i = a.pair_i[t0 + lane]; n_j = a.pair_j[t0 + lane];
                if (EPN) n_near = (float)a.near[t0 + lane];
#pragma unroll
                for (int q = 0; q < EDR / 4; ++q) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(a.e + (int64_t)(t0 + lane) * EDR) + q);
                    cf[4 * q] = x.x; cf[4 * q + 1] = x.y; cf[4 * q + 2] = x.z; cf[4 * q + 3] = x.w;
                }
            }
        };
        if (p0 < p1) fetch_tile(p0);
        for (int tb = p0; tb < p1; tb += 32) {
            const bool ok = tb + lane < p1;
            const int li = n_i - atom0, lj = n_j - atom0;          // slots beyond the tile point at atom 0 (results dropped)
            const float nearf = n_near;
            float ce[HID];
            {
                f2_t ce2[HID / 2];
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) ce2[o] = 0ull;
#pragma unroll
                for (int k = 0; k < EDR; ++k)
#pragma unroll
                    for (int o = 0; o < HID / 2; o += 2) {
                        const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(sCw + k * HID + 2 * o);
                        cfma2(ce2[o], w4.x, cf[k]); cfma2(ce2[o + 1], w4.y, cf[k]);
                    }
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) cunpack2(ce2[o], ce[2 * o], ce[2 * o + 1]);
            }
            if (tb + 32 < p1) fetch_tile(tb + 32);                 // cf is dead from here on
            float fd = 0.f;
#pragma unroll 1
            for (int dir = 0; dir < 2; ++dir) {                    // dir 0: i receives from j;  dir 1: j receives from i
                const int ir = dir ? lj : li, is = dir ? li : lj;
                f2_t acc2[HID / 2];
                second_layer<true>(W, sb2, ce, uv + ir * CUVS, uv + is * CUVS + HID, acc2);
                if (EPN) {
                    float f = 0.f;
#pragma unroll
                    for (int o = 0; o < HID / 2; ++o) {
                        float x, y;
                        cunpack2(acc2[o], x, y);
                        const float2 w3 = *reinterpret_cast<const float2*>(sx + 2 * o);
                        f = fmaf(fmaxf(x, 0.f), w3.x, f);
                        f = fmaf(fmaxf(y, 0.f), w3.y, f);
                    }
                    fd = dir ? fd - f : f;
                } else {
                    add_messages(acc2, 1.f, ok ? ir : -1, Mt, S, lane);
                }
            }
            if (EPN && ok) a.delta[tb + lane] = 0.5f * fd * nearf;          // charge_gn.py:116
        }

        if (!EPN) {
            // ------------------------------------------------------------ far tiles: lane = ordered e == 0 slot (or species slot)
            bool use0 = a.dedup != 0;
            if (use0) {                                            // same check as bundle_kernel: v rows equal species by species?
                bool same = true;
                for (int r = lane; r < nat; r += 32) {
                    const int rp = a.rep[atom0 + r] - atom0;
                    if (rp != r) {
#pragma unroll
                        for (int c = 0; c < HID / 4; ++c) {
                            const float4 x = *reinterpret_cast<const float4*>(uv + r * CUVS + HID + c * 4);
                            const float4 y = *reinterpret_cast<const float4*>(uv + rp * CUVS + HID + c * 4);
                            same = same && x.x == y.x && x.y == y.y && x.z == y.z && x.w == y.w;
                        }
                    }
                }
                use0 = __all_sync(0xffffffffu, same);
            }
            const unsigned short* flist = use0 ? a.far0_list : a.far_list;
            const int f0 = use0 ? a.far0_off[atom0] : a.far_off[atom0], f1 = use0 ? a.far0_off[atom0 + nat] : a.far_off[atom0 + nat];
            const float zero_ce[HID] = {};
            // the next tile's slot codes / weights are requested one tile ahead (two registers)
            int n_code = f0 + lane < f1 ? (int)flist[f0 + lane] : -1;
            int n_cnt = (use0 && f0 + lane < f1) ? (int)a.far0_w[f0 + lane] : 1;
            for (int fb = f0; fb < f1; fb += 32) {
                const bool ok = n_code >= 0;
                int li = 0, lj = 0;
                float wv = 0.f;
                if (ok) {
                    li = n_code >> 8; lj = n_code & 0xFF;
                    if (lj == 0xFF) { lj = PAD_ROW; wv = padw[li]; }        // pad pseudo-pair: weight npad - n, v = b1
                    else wv = (float)n_cnt;                                 // species slot: number of far columns of that species (1: plain column)
                }
                {
                    const int k = fb + 32 + lane;
                    n_code = k < f1 ? (int)flist[k] : -1;
                    n_cnt = (use0 && k < f1) ? (int)a.far0_w[k] : 1;
                }
                f2_t acc2[HID / 2];
                second_layer<false>(W, sb2, zero_ce, uv + li * CUVS, uv + lj * CUVS + HID, acc2);
                add_messages(acc2, wv, ok ? li : -1, Mt, S, lane);
            }
            // ---- S -> global (plane 0 of the partial-sum planes the per-atom kernel reads)
            for (int f = lane; f < nat * (HID / 4); f += 32)
                *reinterpret_cast<float4*>(a.S + (int64_t)atom0 * HID + f * 4) = *reinterpret_cast<const float4*>(S + f * 4);
        }
    }
}

#ifndef EPNN_CPU_EMU
template <bool EPN>
static cudaError_t launch_const(const Workspace& w, const StepW<float>& sw, cudaStream_t st, int* nl) {
    if (w.n_bundles == 0) return cudaSuccess;
    if (!w.wf_host || !w.wf_dev) return cudaErrorInvalidValue;
    PairW W;                                                       // this launch's weights; copied into the launch's parameter buffer
    auto host = [&](const float* dev) { return w.wf_host + (dev - w.wf_dev); };
    memcpy(W.W2, host(sw.W2), sizeof(W.W2));
    ConstArgs ca;
    ca.n_bundles = w.n_bundles; ca.bundle = w.bundle; ca.work_counter = w.work_counter;
    cudaError_t e = cudaMemsetAsync(w.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    ca.ustart = w.ustart; ca.pair_i = w.pair_i; ca.pair_j = w.pair_j; ca.near = w.near; ca.e = w.e;
    ca.far_off = w.far_off; ca.far_list = w.far_list;
    ca.far0_off = w.far0_off; ca.far0_list = w.far0_list; ca.far0_w = w.far0_w; ca.rep = w.rep; ca.dedup = w.dedup_far;
    ca.atom_sys = w.atom_sys; ca.sys_off = w.sys_off; ca.npad = w.npad;
    ca.u = (const float*)w.u; ca.v = (const float*)w.v; ca.S = (float*)w.S; ca.delta = (float*)w.delta;
    ca.Cw = sw.Cw; ca.b2 = sw.b2; ca.x32 = EPN ? sw.W3 : sw.b1;
    const size_t smem = ConstSmem<EPN>::bytes();
    e = cudaFuncSetAttribute(bundle_const_kernel<EPN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_sm = EPN ? CONST_CTAS_EPN : 1;
    int grid = div_up(w.n_bundles, CONST_NW);
    if (grid > w.sm_count * per_sm) grid = w.sm_count * per_sm;
    e = cudaMemcpyAsync(w.args_dev, &ca, sizeof(ca), cudaMemcpyHostToDevice, st);      // pageable source: staged before the call returns
    if (e != cudaSuccess) return e;
    bundle_const_kernel<EPN><<<grid, CONST_NW * 32, smem, st>>>(W, (const ConstArgs*)w.args_dev);
    ++*nl;
    return cudaGetLastError();
}

cudaError_t launch_gnn_bundle_const(const Workspace& w, const StepW<float>& sw, cudaStream_t st, int* nl) { return launch_const<false>(w, sw, st, nl); }
cudaError_t launch_epn_bundle_const(const Workspace& w, const StepW<float>& sw, cudaStream_t st, int* nl) { return launch_const<true>(w, sw, st, nl); }
#endif   // !EPNN_CPU_EMU
