// EXPERIMENTAL FP32 variant of the per-atom kernel (option "pair_const" = 1, together with epnn_bundle_const.cu;
// precision 32 only, default OFF).  Like the bundle kernels of that option it has NOT run on a GPU yet: it compiles for
// sm_100a, and its logic is checked by the CPU warp emulation (tests/test_emu_atom_epn.py runs it next to the default
// atom_kernel on the same inputs); its GPU tests are the gated ones of tests/test_gpu_pair_const.py.
//
// Same modes and the same arithmetic as atom_kernel<float, NW> (epnn_atom.cu), other mapping: ONE THREAD OWNS ONE ATOM.
// The three products of a message-passing step -- [l2_prev | S] (64) -> l1 (32) -> l2 (32) -> u | v (64), 5 120 MAC --
// keep their activations and accumulators in the thread's registers and take the weights as uniform operands from a
// __grid_constant__ kernel parameter (FFMA2 R, R.F32, UR.F32x2, R.F32x2; see epnn_bundle_const.cu), so there is no
// shared-memory operand traffic.  Shared memory is only a transpose buffer: a warp's 32 rows are read from / written to
// global memory as whole 128-byte lines and handed to / taken from the owning lanes through a padded tile.
#include "epnn_internal.cuh"              // (under EPNN_CPU_EMU this pulls in the CPU warp-emulation shim)

#define ACONST_NW 8
#define ATS 68                            // row stride of the transpose tile: 64 + 4 floats

struct alignas(16) AtomW {      // (read with 8-byte loads: keep every member offset a multiple of 8)
    float HG[64 * HID]; float cb[HID]; float g[HID];          // first update layer on [l2_prev | S], its bias, U1_M^T b3 (times npad)
    float U2[HID * HID]; float c2[HID];                        // second update layer
    float U3[HID * HD]; float c3[HD];                          // h = U3^T l2 + c3 (last step only)
    float Pf[HID * 64]; float Aq[64];                          // projections of the next pair kernel: U3 . Ah64, q row
};

static_assert(sizeof(AtomW) % 16 == 0 && sizeof(AtomW) < 32000, "AtomW is a __grid_constant__ kernel parameter: 32 764 bytes at most (CUDA >= 12.1)");

struct AtomConstArgs {
    int n_atoms, mode, nsplit, h_is_zero;
    const int* atom_sys; const int* sys_off; const int* npad; const int* species;
    const float* Spart; float* h; float* l2;
    const int* rowptr; const int* col; const int* pid; const float* delta; double* q;
    const float* Ax;                                           // [MAX_SPECIES][64] per-species table (not uniform: gathered per lane)
    float* u; float* v; float* q_out; double* q_out64;
};

typedef unsigned long long a2_t;
#ifdef EPNN_CPU_EMU
__device__ __forceinline__ a2_t apack2(float lo, float hi) { unsigned a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return (a2_t)a | ((a2_t)b << 32); }
__device__ __forceinline__ void aunpack2(a2_t v, float& lo, float& hi) { const unsigned a = (unsigned)v, b = (unsigned)(v >> 32); memcpy(&lo, &a, 4); memcpy(&hi, &b, 4); }
__device__ __forceinline__ void afma2(a2_t& d, a2_t wpair, float a) {
    float d0, d1, w0, w1;
    aunpack2(d, d0, d1); aunpack2(wpair, w0, w1);
    d = apack2(fmaf(w0, a, d0), fmaf(w1, a, d1));
}
#else
__device__ __forceinline__ a2_t apack2(float lo, float hi) { a2_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void aunpack2(a2_t v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void afma2(a2_t& d, a2_t wpair, float a) {
    const a2_t aa = apack2(a, a);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(wpair), "l"(aa));
}
#endif

// out[0 .. 2*NP) += x[0 .. K) . Wm[K][ld] (columns col0 .. col0 + 2*NP), weights uniform.  Fully unrolled on purpose: the
// activations x[] live in registers, and a rolled k loop would index them dynamically (ptxas then moves them to local
// memory: tried, 256-byte stack frame).  The price is code size -- 4 864 FFMA2, about 120 KB of SASS for the kernel.
template <int K, int NP>
__device__ __forceinline__ void uniform_product(const float* __restrict__ Wm, int ld, int col0, const float (&x)[K], a2_t (&acc)[NP]) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int o = 0; o < NP; ++o) afma2(acc[o], *reinterpret_cast<const a2_t*>(&Wm[k * ld + col0 + 2 * o]), x[k]);
}

__global__ void __launch_bounds__(ACONST_NW * 32, 2) atom_const_kernel(const __grid_constant__ AtomW W, const AtomConstArgs a) {
#ifdef EPNN_CPU_EMU
    float* asm_ = emu_smem;
#else
    extern __shared__ __align__(16) float asm_[];
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* T = asm_ + warp * (32 * ATS);                       // [32][ATS] transpose tile of this warp
    const bool do_upd = a.mode & ATOM_UPDATE, do_q = a.mode & ATOM_QUPDATE, do_proj = a.mode & ATOM_PROJECT;
    const bool first = a.mode & ATOM_FIRST, write_h = a.mode & ATOM_WRITE_H;
    const int n_tiles = (a.n_atoms + 31) / 32;

    for (int tile = blockIdx.x * ACONST_NW + warp; tile < n_tiles; tile += gridDim.x * ACONST_NW) {
        const int base = tile * 32;
        const int me = base + lane;
        const bool me_ok = me < a.n_atoms;
        // ---------------- per-atom scalars and the charge update (lane = atom)
        int sp = 0, ns = 1;
        float npf = 0.f;
        double qv = 0.0;
        if (me_ok) {
            const int sys = a.atom_sys[me];
            sp = a.species[me];
            ns = a.sys_off[sys + 1] - a.sys_off[sys] > SMALL_MAX ? a.nsplit : 1;
            npf = (float)a.npad[sys];
            qv = a.q[me];
            if (do_q) {
                const int r0 = a.rowptr[me], r1 = a.rowptr[me + 1];
                for (int k = r0; k < r1; ++k) {                // fixed (ascending column) order, FP64: deterministic, conserving
                    const double d = (double)a.delta[a.pid[k]];
                    qv += a.col[k] > me ? d : -d;
                }
                a.q[me] = qv;
            }
            if (a.mode & ATOM_OUTPUT) {
                if (a.q_out) a.q_out[me] = (float)qv;
                if (a.q_out64) a.q_out64[me] = qv;
            }
        }
        if (!do_upd && !do_proj) continue;

        float l2v[HID];                                        // this atom's l2 (update output, or read back for the EPN projections)
        if (do_upd) {
            // (1) [l2_prev | S] rows: coalesced loads (lane = 16-byte chunk), summed planes, into the transpose tile
            __syncwarp();
#pragma unroll 2
            for (int f = lane; f < 32 * 8; f += 32) {
                const int sl = f >> 3, ch = f & 7;
                const int at = base + sl;
                float4 lv = make_float4(0.f, 0.f, 0.f, 0.f), sv = make_float4(0.f, 0.f, 0.f, 0.f);
                const int nsl = __shfl_sync(0xffffffffu, ns, sl);       // partial-sum planes of that row's system (all lanes take part)
                if (at < a.n_atoms) {
                    if (!first) lv = *reinterpret_cast<const float4*>(a.l2 + (int64_t)at * HID + ch * 4);
                    for (int p = 0; p < nsl; ++p) {                     // planes added in fixed order
                        const float4 x = *reinterpret_cast<const float4*>(a.Spart + ((int64_t)p * a.n_atoms + at) * HID + ch * 4);
                        sv.x += x.x; sv.y += x.y; sv.z += x.z; sv.w += x.w;
                    }
                }
                *reinterpret_cast<float4*>(T + sl * ATS + ch * 4) = lv;
                *reinterpret_cast<float4*>(T + sl * ATS + HID + ch * 4) = sv;
            }
            __syncwarp();
            // (2) l1 = relu([U3 U1_h ; W3 U1_M]^T [l2_prev | S] + cb + npad * g)
            float l1[HID];
            {
                a2_t acc[HID / 2];
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) acc[o] = apack2(fmaf(npf, W.g[2 * o], W.cb[2 * o]), fmaf(npf, W.g[2 * o + 1], W.cb[2 * o + 1]));
#pragma unroll
                for (int k8 = 0; k8 < 8; ++k8) {               // 8 inputs at a time: two 16-byte reads of the lane's own row
                    float x[8];
                    const float4 x0 = *reinterpret_cast<const float4*>(T + lane * ATS + 8 * k8);
                    const float4 x1 = *reinterpret_cast<const float4*>(T + lane * ATS + 8 * k8 + 4);
                    x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
                    uniform_product<8, HID / 2>(W.HG + 8 * k8 * HID, HID, 0, x, acc);
                }
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) { float p, r; aunpack2(acc[o], p, r); l1[2 * o] = fmaxf(p, 0.f); l1[2 * o + 1] = fmaxf(r, 0.f); }
            }
            // (3) l2 = relu(U2^T l1 + c2)
            {
                a2_t acc[HID / 2];
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) acc[o] = apack2(W.c2[2 * o], W.c2[2 * o + 1]);
                uniform_product<HID, HID / 2>(W.U2, HID, 0, l1, acc);
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) { float p, r; aunpack2(acc[o], p, r); l2v[2 * o] = fmaxf(p, 0.f); l2v[2 * o + 1] = fmaxf(r, 0.f); }
            }
            // l2 -> global through the tile (whole lines)
            __syncwarp();
#pragma unroll
            for (int c4 = 0; c4 < HID / 4; ++c4)
                *reinterpret_cast<float4*>(T + lane * ATS + c4 * 4) = make_float4(l2v[4 * c4], l2v[4 * c4 + 1], l2v[4 * c4 + 2], l2v[4 * c4 + 3]);
            __syncwarp();
#pragma unroll 2
            for (int f = lane; f < 32 * 8; f += 32) {
                const int sl = f >> 3, ch = f & 7;
                if (base + sl < a.n_atoms) *reinterpret_cast<float4*>(a.l2 + (int64_t)(base + sl) * HID + ch * 4) = *reinterpret_cast<const float4*>(T + sl * ATS + ch * 4);
            }
            // (4) last message-passing step: h = U3^T l2 + c3 (48 columns), straight from the owning lane
            if (write_h) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {         // columns 0..23, 24..47
                    a2_t acc[12];
#pragma unroll
                    for (int o = 0; o < 12; ++o) acc[o] = apack2(W.c3[24 * half + 2 * o], W.c3[24 * half + 2 * o + 1]);
                    uniform_product<HID, 12>(W.U3, HD, 24 * half, l2v, acc);
                    if (me_ok) {
#pragma unroll
                        for (int o = 0; o < 12; o += 2) {
                            float p0, p1, p2, p3;
                            aunpack2(acc[o], p0, p1); aunpack2(acc[o + 1], p2, p3);
                            *reinterpret_cast<float4*>(a.h + (int64_t)me * HD + 24 * half + 2 * o) = make_float4(p0, p1, p2, p3);
                        }
                    }
                }
            }
        } else if (do_proj && !a.h_is_zero) {
            // l2 of the last message-passing step, rows through the tile
            __syncwarp();
#pragma unroll 2
            for (int f = lane; f < 32 * 8; f += 32) {
                const int sl = f >> 3, ch = f & 7;
                float4 lv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + sl < a.n_atoms) lv = *reinterpret_cast<const float4*>(a.l2 + (int64_t)(base + sl) * HID + ch * 4);
                *reinterpret_cast<float4*>(T + sl * ATS + ch * 4) = lv;
            }
            __syncwarp();
#pragma unroll
            for (int c4 = 0; c4 < HID / 4; ++c4) {
                const float4 x = *reinterpret_cast<const float4*>(T + lane * ATS + c4 * 4);
                l2v[4 * c4] = x.x; l2v[4 * c4 + 1] = x.y; l2v[4 * c4 + 2] = x.z; l2v[4 * c4 + 3] = x.w;
            }
        }

        if (do_proj) {
            // u | v = (U3 Ah)^T l2 + (Ax[species] + c3^T Ah) + q Aq      (first step: h = 0, no product)
            const float qf = (float)qv;
            __syncwarp();
#pragma unroll
            for (int half = 0; half < 2; ++half) {             // u (columns 0..31), then v (32..63)
                a2_t acc[HID / 2];
#pragma unroll
                for (int o = 0; o < HID / 2; ++o) acc[o] = 0ull;
                if (!a.h_is_zero) uniform_product<HID, HID / 2>(W.Pf, 64, HID * half, l2v, acc);
#pragma unroll
                for (int c4 = 0; c4 < HID / 4; ++c4) {
                    const float4 ax = *reinterpret_cast<const float4*>(a.Ax + sp * 64 + HID * half + c4 * 4);
                    float p0, p1, p2, p3;
                    aunpack2(acc[2 * c4], p0, p1); aunpack2(acc[2 * c4 + 1], p2, p3);
                    float4 o4;
                    o4.x = p0 + fmaf(qf, W.Aq[HID * half + 4 * c4], ax.x); o4.y = p1 + fmaf(qf, W.Aq[HID * half + 4 * c4 + 1], ax.y);
                    o4.z = p2 + fmaf(qf, W.Aq[HID * half + 4 * c4 + 2], ax.z); o4.w = p3 + fmaf(qf, W.Aq[HID * half + 4 * c4 + 3], ax.w);
                    *reinterpret_cast<float4*>(T + lane * ATS + HID * half + c4 * 4) = o4;
                }
            }
            __syncwarp();
#pragma unroll 2
            for (int f = lane; f < 32 * 16; f += 32) {         // whole lines out: u rows, v rows
                const int sl = f >> 4, ch = f & 15;
                if (base + sl < a.n_atoms) {
                    float* dst = (ch < 8 ? a.u : a.v) + (int64_t)(base + sl) * HID + (ch & 7) * 4;
                    *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(T + sl * ATS + ch * 4);
                }
            }
        }
        __syncwarp();
    }
}

#ifndef EPNN_CPU_EMU
cudaError_t launch_atom_const(const Workspace& w, int mode, const StepW<float>* prev, const UpdW<float>* upd, const StepW<float>* next,
                              int h_is_zero, float* q_out, double* q_out64, cudaStream_t st, int* nl) {
    if (w.n_atoms == 0) return cudaSuccess;
    if (!w.wf_host || !w.wf_dev) return cudaErrorInvalidValue;
    auto host = [&](const float* dev) { return w.wf_host + (dev - w.wf_dev); };
    AtomW W;
    memset(&W, 0, sizeof(W));
    AtomConstArgs aa;
    memset(&aa, 0, sizeof(aa));
    if (mode & ATOM_UPDATE) {
        memcpy(W.HG, host(prev->HG), sizeof(W.HG)); memcpy(W.g, host(prev->g), sizeof(W.g));
        memcpy(W.cb, host((mode & ATOM_FIRST) ? upd->c1 : upd->cb1), sizeof(W.cb));
        memcpy(W.U2, host(upd->U2), sizeof(W.U2)); memcpy(W.c2, host(upd->c2), sizeof(W.c2));
        memcpy(W.U3, host(upd->U3), sizeof(W.U3)); memcpy(W.c3, host(upd->c3), sizeof(W.c3));
    }
    if (mode & ATOM_PROJECT) {
        memcpy(W.Pf, host(next->Pf), sizeof(W.Pf)); memcpy(W.Aq, host(next->Aq64), sizeof(W.Aq));
        aa.Ax = h_is_zero ? next->Ax64 : next->Axf;
    }
    aa.n_atoms = w.n_atoms; aa.mode = mode; aa.nsplit = w.nsplit; aa.h_is_zero = h_is_zero;
    aa.atom_sys = w.atom_sys; aa.sys_off = w.sys_off; aa.npad = w.npad; aa.species = w.species;
    aa.Spart = (const float*)w.S; aa.h = (float*)w.h; aa.l2 = (float*)w.l2;
    aa.rowptr = w.rowptr; aa.col = w.col; aa.pid = w.pid; aa.delta = (const float*)w.delta; aa.q = w.q;
    aa.u = (float*)w.u; aa.v = (float*)w.v; aa.q_out = q_out; aa.q_out64 = q_out64;
    const size_t smem = sizeof(float) * ACONST_NW * 32 * ATS;
    cudaError_t e = cudaFuncSetAttribute(atom_const_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = div_up(div_up(w.n_atoms, 32), ACONST_NW);
    if (grid > 2 * w.sm_count) grid = 2 * w.sm_count;
    atom_const_kernel<<<grid, ACONST_NW * 32, smem, st>>>(W, aa);
    ++*nl;
    return cudaGetLastError();
}
#endif
