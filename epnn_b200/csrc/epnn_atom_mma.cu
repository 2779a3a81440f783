// Per-atom kernel on the warp-level tensor path (FP32 calls; option "atom_tensor", default on).
//
// Same modes, inputs and outputs as atom_kernel<float, NW, float> (epnn_atom.cu; reference charge_gn.py:70-74, :116-118 and
// the first-layer projections of :62-68 / :101-110), but the three dense layers of a step
//     [l2_prev | S] (64) -> l1 (32) -> l2 (32) -> u | v (64)      (+ h = U3^T l2 + c3 (48) at the last message-passing step)
// are real GEMMs over all atoms of the chunk (M = atoms, K / N = 32 .. 64), so they run as mma.sync.m16n8k8 (TF32 inputs,
// FP32 accumulation) on the two-term split  x = hi + lo  of both operands, hi and lo rounded to TF32 (nearest, ties away):
// (hi + lo)(hi + lo) = hi*hi + lo*hi + hi*lo + lo*lo, FOUR MMAs per block.  The usual 3xTF32 scheme drops lo*lo (2^-22 of a
// product) and lets the tensor core truncate lo; measured over all 4 374 systems of data/mixed with model2_weights (pad 41,
// profiles/r02/call63_*, call64_*) that left a tail 1.5x the FP32 SIMT kernel's -- one system at 1.12e-5 e, above north_star's
// 1e-5 -- while the full product with rounded lo is at 9.2e-6 / p99.9 4.7e-6 against 7.9e-6 / 4.0e-6 for SIMT, for 0.7 ms
// more per 300 k molecules.  The three correction terms chain in their own accumulator (2^-11 of the main term: the tensor
// core's round-toward-zero is irrelevant there), the hi*hi blocks of one am_pair call chain for at most four k blocks and
// are added on the FP32 pipe (round to nearest).  With the FP32 pipe out of the way the kernel is bound by its HBM traffic
// (640 B per atom and step) and by instruction issue, not by shared-memory operand reads.
//
// A warp owns a tile of 32 consecutive atoms = two 16-row m-tiles; thread (g = lane >> 2, t = lane & 3) owns rows
// g, g + 8, g + 16, g + 24.  ONE index map serves every operand,
//     amap(blk, t, h) = 16 (blk >> 1) + 4 t + 2 (blk & 1) + h          blk = 8-wide k block or n tile, h = 0 / 1
// fragment position t (h = 0) / t + 4 (h = 1) of k block blk, and C-fragment column 2t + h of n tile blk, both stand for
// the actual feature index amap(blk, t, h).  Consequences: (1) what a thread holds of blocks 2m, 2m + 1 is the float4 at
// columns 16m + 4t .. + 3 of its rows, so every global load and store is a 128-bit access and a warp instruction touches
// 8 rows x 64 contiguous bytes; (2) the C fragments a layer leaves behind ARE the A fragments of the next layer (after bias,
// ReLU and the split) -- no shared-memory stage between layers; (3) the weights are permuted once per CTA into B fragments in
// shared memory: fragment (kb, nt) = 32 lanes x {b0 hi, b1 hi, b0 lo, b1 lo}, one conflict-free LDS.128 per lane feeds eight MMAs.
#include "epnn_internal.cuh"

#ifndef AM_CHAIN
#define AM_CHAIN 1                               // 1: the hi*hi blocks of one am_pair call chain inside the tensor core; 0: each block is added on the FP32 pipe
#endif
#ifndef AM_LOLO
#define AM_LOLO 1                                // fourth MMA lo*lo (2^-22 of the product): see the header, 0 only for A/B
#endif
#ifndef AM_ROUND_LO
#define AM_ROUND_LO 1                            // lo rounded to TF32 (nearest) instead of truncated by the tensor core; 0 only for A/B
#endif
#define AM_FRAG 128                              // words per B fragment (32 lanes x 4)
// Shared-memory layout (32-bit words) of the two instantiations: UPD = launches that finish a message-passing step (all three
// layers, 6 warps x 168 registers, 2 CTAs per SM), !UPD = the projections between two electron-passing passes (one layer,
// 8 warps x 128 registers, 2 CTAs per SM: the serial CSR walk of the charge update wants the warps).
template <bool UPD> struct AmL {
    static constexpr int NW = UPD ? 6 : 8;
    static constexpr int B1 = 0;                                       // [8 kb][4 nt]  first update layer  [U3 U1_h ; W3 U1_M]
    static constexpr int B2 = B1 + (UPD ? 32 * AM_FRAG : 0);           // [4][4]        second update layer U2
    static constexpr int B3 = B2 + (UPD ? 16 * AM_FRAG : 0);           // [4][8]        projections of the next pair kernel, U3 Ah64 (u | v)
    static constexpr int BH = B3 + 32 * AM_FRAG;                       // [4][6]        h = U3^T l2 (last message-passing step)
    static constexpr int VEC = BH + (UPD ? 24 * AM_FRAG : 0);          // cb[32] g[32] c2[32] c3[48] aq[64]
    static constexpr int AX = VEC + 208;                               // [MAX_SPECIES][64]
    static constexpr int SLOT = AX + MAX_SPECIES * 64;                 // per warp: q[32] np[32] sp[32] ns[32]
    static constexpr int BAR = SLOT + NW * 128;                        // per warp: one mbarrier (8 bytes) -- "this warp's input tile has landed"
    static constexpr int STAGE = (BAR + 2 * NW + 31) / 32 * 32;        // per warp (128-byte aligned): the tile's l2 rows [32][32] (| plane 0 of
    static constexpr int STAGE_W = UPD ? 2048 : 1024;                  //   its S rows [32][32]), filled by bulk copies
    static constexpr int WORDS = STAGE + NW * STAGE_W;
};

#ifdef EPNN_CPU_EMU
__device__ __forceinline__ void am_mma(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    const unsigned b[2] = {b0, b1};
    emu_mma_m16n8k8_tf32(d, a, b);
}
struct am_u4 { unsigned x, y, z, w; };
// bulk-copy shim: the issuing lane copies at once; the consumers' "wait" is the warp barrier that follows
__device__ __forceinline__ void am_bar_init(void*) {}
__device__ __forceinline__ void am_bar_expect(void*, unsigned) {}
__device__ __forceinline__ void am_bulk(void* dst, const void* src, unsigned bytes, void*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void am_bar_wait(void*, unsigned) { __syncwarp(); }
#else
__device__ __forceinline__ void am_mma(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
typedef uint4 am_u4;
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (one elected lane issues, all lanes wait)
__device__ __forceinline__ unsigned am_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void am_bar_init(void* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(am_saddr(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void am_bar_expect(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(am_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void am_bulk(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(am_saddr(dst)), "l"(src), "r"(bytes), "r"(am_saddr(bar)) : "memory");
}
__device__ __forceinline__ void am_bar_wait(void* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(am_saddr(bar)), "r"(parity) : "memory");
    } while (!ok);
}
#endif

// x = hi + lo: hi = x rounded to TF32 (nearest, ties away from zero: integer add on the magnitude bits), lo = x - hi (exact)
// rounded to TF32 the same way -- the tensor core would truncate it
__device__ __forceinline__ void am_split(float x, unsigned& hi, unsigned& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
#if AM_ROUND_LO
    lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xFFFFE000u;
#else
    lo = __float_as_uint(x - __uint_as_float(hi));
#endif
}
__device__ __forceinline__ int am_map(int blk, int t, int h) { return 16 * (blk >> 1) + 4 * t + 2 * (blk & 1) + h; }

// B fragments of W[K][ld] columns col0 .. col0 + 8 NT: dst[(kb * NT + nt) * 128 + lane * 4] = {b0 hi, b1 hi, b0 lo, b1 lo},
// b0 = W[amap(kb, t, 0)][col0 + amap(nt, g >> 1, g & 1)], b1 = W[amap(kb, t, 1)][same column].
__device__ __forceinline__ void am_stage(unsigned* dst, const float* __restrict__ W, int ld, int col0, int KB, int NT, int tid, int nthr) {
    for (int f = tid; f < KB * NT * 32; f += nthr) {
        const int ln = f & 31, fr = f >> 5, kb = fr / NT, nt = fr - kb * NT, g = ln >> 2, t = ln & 3;
        const int n = col0 + am_map(nt, g >> 1, g & 1);
        unsigned h0, l0, h1, l1;
        am_split(W[am_map(kb, t, 0) * ld + n], h0, l0);
        am_split(W[am_map(kb, t, 1) * ld + n], h1, l1);
        unsigned* d = dst + (size_t)f * 4;
        d[0] = h0; d[1] = h1; d[2] = l0; d[3] = l1;
    }
}

// Fragment sets are stored per column pair: z[m][mt][n][e] = C-fragment element e (d0 (row, 2t) d1 (row, 2t + 1) d2 (row + 8, 2t)
// d3 (row + 8, 2t + 1)) of m-tile mt and n tile 2m + n.
//
// acc[mt][n] += A (two m-tiles, KB k blocks, hi / lo) * B fragments (kb0 + kb, nt0 + n) for one column pair.  The tensor core
// rounds its FP32 accumulator toward zero, a bias that grows with the length of an accumulation chain (first GPU run of this
// kernel, everything in one chain: model2_weights 3.9e-6 -> 9.8e-6 from the oracle).  So the small correction terms chain in
// their own accumulator, and the hi*hi blocks either chain for the k blocks of this one call (AM_CHAIN, default) or -- as in
// Ootomo & Yokota's error-corrected TF32 GEMM -- go into a ZERO accumulator each and are added on the FP32 pipe (measured: the
// same noise, 600 more FADDs per tile).
template <int KB, bool SET = false>        // SET: acc = product (acc need not be initialised), else acc += product
__device__ __forceinline__ void am_pair(const unsigned (&ah)[2][KB][4], const unsigned (&al)[2][KB][4], const unsigned* __restrict__ sB,
                                        int kb0, int ntot, int nt0, int lane, float (&acc)[2][2][4]) {
    float corr[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            corr[mt][n][0] = corr[mt][n][1] = corr[mt][n][2] = corr[mt][n][3] = 0.f;
#if !AM_CHAIN
            if (SET) acc[mt][n][0] = acc[mt][n][1] = acc[mt][n][2] = acc[mt][n][3] = 0.f;
#endif
        }
#if AM_CHAIN
    float mainacc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 2; ++n) mainacc[mt][n][0] = mainacc[mt][n][1] = mainacc[mt][n][2] = mainacc[mt][n][3] = 0.f;
#endif
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const am_u4 b = *reinterpret_cast<const am_u4*>(sB + ((size_t)((kb0 + kb) * ntot + nt0 + n) * 32 + lane) * 4);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                am_mma(corr[mt][n], al[mt][kb], b.x, b.y);
                am_mma(corr[mt][n], ah[mt][kb], b.z, b.w);
#if AM_LOLO
                am_mma(corr[mt][n], al[mt][kb], b.z, b.w);
#endif
#if AM_CHAIN
                am_mma(mainacc[mt][n], ah[mt][kb], b.x, b.y);
#else
                float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                am_mma(tmp, ah[mt][kb], b.x, b.y);
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[mt][n][e] += tmp[e];
#endif
            }
        }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#if AM_CHAIN
                if (SET) acc[mt][n][e] = mainacc[mt][n][e] + corr[mt][n][e];
                else acc[mt][n][e] += mainacc[mt][n][e] + corr[mt][n][e];
#else
                acc[mt][n][e] += corr[mt][n][e];
#endif
            }
}

// C-fragment values of NP column pairs -> A fragments of the next layer: k block 2m + n, a0 = d0, a1 = d2, a2 = d1, a3 = d3
template <int NP>
__device__ __forceinline__ void am_to_a(const float (&z)[NP][2][2][4], unsigned (&ah)[2][2 * NP][4], unsigned (&al)[2][2 * NP][4]) {
#pragma unroll
    for (int m = 0; m < NP; ++m)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                am_split(z[m][mt][n][0], ah[mt][2 * m + n][0], al[mt][2 * m + n][0]);
                am_split(z[m][mt][n][2], ah[mt][2 * m + n][1], al[mt][2 * m + n][1]);
                am_split(z[m][mt][n][1], ah[mt][2 * m + n][2], al[mt][2 * m + n][2]);
                am_split(z[m][mt][n][3], ah[mt][2 * m + n][3], al[mt][2 * m + n][3]);
            }
}
__device__ __forceinline__ void am_zero(float (&acc)[2][2][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 2; ++n) acc[mt][n][0] = acc[mt][n][1] = acc[mt][n][2] = acc[mt][n][3] = 0.f;
}
// row i (0..3 <-> g + 8 i) of one column pair as the float4 at columns 16 m + 4 t .. + 3
__device__ __forceinline__ float4 am_get4(const float (&z)[2][2][4], int i) {
    const int mt = i >> 1, e = (i & 1) * 2;
    return make_float4(z[mt][0][e], z[mt][0][e + 1], z[mt][1][e], z[mt][1][e + 1]);
}
__device__ __forceinline__ void am_put4(float (&z)[2][2][4], int i, float4 v) {
    const int mt = i >> 1, e = (i & 1) * 2;
    z[mt][0][e] = v.x; z[mt][0][e + 1] = v.y; z[mt][1][e] = v.z; z[mt][1][e + 1] = v.w;
}
__device__ __forceinline__ float4 am_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void am_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float am_relu(float x) { return x > 0.f ? x : 0.f; }

template <bool SCOPED, bool UPD>
__global__ void __launch_bounds__(AmL<UPD>::NW * 32, 2) atom_mma_kernel(const AtomArgs<float, float> a) {
    typedef AmL<UPD> L;
    constexpr int AM_NW = L::NW;
#ifdef EPNN_CPU_EMU
    unsigned* sm = reinterpret_cast<unsigned*>(emu_smem);
#else
    extern __shared__ __align__(16) unsigned sm[];
#endif
    float* sv = reinterpret_cast<float*>(sm + L::VEC);
    float* scb = sv, *sg = sv + 32, *sc2 = sv + 64, *sc3 = sv + 96, *saq = sv + 144;
    float* sAx = reinterpret_cast<float*>(sm + L::AX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float* slot_q = reinterpret_cast<float*>(sm + L::SLOT + warp * 128);
    float* slot_np = slot_q + 32;
    int* slot_sp = reinterpret_cast<int*>(slot_np + 32);
    int* slot_ns = slot_sp + 32;

    constexpr bool do_upd = UPD;                         // == (a.mode & ATOM_UPDATE): the launcher picks the instantiation
    const bool do_q = a.mode & ATOM_QUPDATE, do_proj = a.mode & ATOM_PROJECT;
    const bool first = a.mode & ATOM_FIRST, write_h = a.mode & ATOM_WRITE_H;
    const bool proj_gemm = do_proj && !a.h_is_zero;
    const int nthr = AM_NW * 32;
    // input pipeline: the warp's NEXT tile (l2 rows and / or plane 0 of the S rows: 4 KB each, contiguous in global memory) is
    // fetched by two bulk copies as soon as the current tile's inputs have been read out of the stage
    float* st_l2 = reinterpret_cast<float*>(sm + L::STAGE + warp * L::STAGE_W);
    float* st_S = st_l2 + 1024;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + L::BAR) + warp;
    const bool need_l2 = (do_upd && !first) || (proj_gemm && !do_upd), need_S = do_upd;
    const int n_tiles = (a.n_atoms + 31) / 32;
    auto fetch = [&](int tile) {                         // one lane
        const int rows = min(32, a.n_atoms - tile * 32);
        const unsigned bytes = (unsigned)rows * HID * sizeof(float);
        am_bar_expect(bar, bytes * ((need_l2 ? 1u : 0u) + (need_S ? 1u : 0u)));
        if (need_l2) am_bulk(st_l2, a.l2 + (int64_t)tile * 32 * HID, bytes, bar);
        if (need_S) am_bulk(st_S, a.Spart + (int64_t)tile * 32 * HID, bytes, bar);
    };
    if (lane == 0) am_bar_init(bar);
    if (do_upd) {
        am_stage(sm + L::B1, a.HG, HID, 0, 8, 4, threadIdx.x, nthr);
        am_stage(sm + L::B2, a.upd.U2, HID, 0, 4, 4, threadIdx.x, nthr);
        if (write_h) am_stage(sm + L::BH, a.upd.U3, HD, 0, 4, 6, threadIdx.x, nthr);
        if (threadIdx.x < HID) { scb[threadIdx.x] = a.cb[threadIdx.x]; sg[threadIdx.x] = a.g[threadIdx.x]; sc2[threadIdx.x] = a.upd.c2[threadIdx.x]; }
        if (threadIdx.x < HD) sc3[threadIdx.x] = a.upd.c3[threadIdx.x];
    }
    if (do_proj) {
        if (proj_gemm) am_stage(sm + L::B3, a.Pf, 64, 0, 4, 8, threadIdx.x, nthr);
        for (int f = threadIdx.x; f < MAX_SPECIES * 64; f += nthr) sAx[f] = a.Ax[f];
        if (threadIdx.x < 64) saq[threadIdx.x] = a.Aq64[threadIdx.x];
    }
    __syncthreads();

    const int tile0 = blockIdx.x * AM_NW + warp, tstep = gridDim.x * AM_NW;
    if (lane == 0 && tile0 < n_tiles) fetch(tile0);
    unsigned parity = 0;
    for (int tile = tile0; tile < n_tiles; tile += tstep) {
        const int base = tile * 32;
        const int me = base + lane;
        const bool me_ok = me < a.n_atoms;
        const bool more = tile + tstep < n_tiles;
        // ---------------- per-slot scalars and the charge update (lane = slot)
        {
            int sp = 0, ns = 0; float npf = 0.f;         // ns = 0 marks a slot this launch does not touch
            double qv = 0.0;
            bool in = me_ok;
            int sys = 0, nat = 0;
            if (me_ok) {
                sys = a.atom_sys[me];
                nat = a.sys_off[sys + 1] - a.sys_off[sys];
                if (SCOPED && a.scope && nat > SMALL_MAX) in = a.scope == 1 ? (me >= a.row_lo && me < a.row_hi) : a.active[me] != 0;
            }
            if (in) {
                sp = a.species[me];
                ns = nat > SMALL_MAX ? a.nsplit : 1;
                npf = (float)a.npad[sys];
                qv = a.q[me];
                if (do_q) {
                    const int r0 = a.rowptr[me], r1 = a.rowptr[me + 1];
                    for (int k = r0; k < r1; ++k) {          // fixed (ascending column) order
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
            slot_sp[lane] = sp; slot_ns[lane] = ns; slot_np[lane] = npf; slot_q[lane] = (float)qv;
        }
        __syncwarp();
        int rns[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rns[i] = slot_ns[g + 8 * i];
        am_bar_wait(bar, parity);                        // the tile's inputs are in the stage
        parity ^= 1u;
        if (SCOPED && a.scope && !__any_sync(0xffffffffu, rns[0] | rns[1] | rns[2] | rns[3])) {      // nothing of this tile belongs to the launch
            __syncwarp();
            if (lane == 0 && more) fetch(tile + tstep);
            continue;
        }

        unsigned ah[2][4][4], al[2][4][4];               // A fragments of the layer about to run (k blocks 0..3)
        float z[2][2][2][4];                             // a layer's 32 output columns: two column pairs
        if (do_upd) {
            // ---- first layer: relu([U3 U1_h ; W3 U1_M]^T [l2_prev | S] + cb + npad g), 32 input columns (four k blocks) at a time:
            // the l2 half (skipped at the first step: h = 0), then the S half
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                if (jj == 0 && first) continue;
                float zz[2][2][2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rns[i] > 0) {
                            if (jj == 0) x = am_ld4(st_l2 + (g + 8 * i) * HID + 16 * m + 4 * t);
                            else {
                                x = am_ld4(st_S + (g + 8 * i) * HID + 16 * m + 4 * t);
                                for (int sp = 1; sp < rns[i]; ++sp) {       // large systems: further partial planes, summed in fixed order
                                    const float4 p = am_ld4(a.Spart + ((int64_t)sp * a.n_atoms + base + g + 8 * i) * HID + 16 * m + 4 * t);
                                    x.x += p.x; x.y += p.y; x.z += p.z; x.w += p.w;
                                }
                            }
                        }
                        am_put4(zz[m], i, x);
                    }
                am_to_a<2>(zz, ah, al);
                if (jj == 0 || first) {
#pragma unroll
                    for (int m = 0; m < 2; ++m) am_pair<4, true>(ah, al, sm + L::B1, 4 * jj, 4, 2 * m, lane, z[m]);
                } else {
#pragma unroll
                    for (int m = 0; m < 2; ++m) am_pair<4, false>(ah, al, sm + L::B1, 4 * jj, 4, 2 * m, lane, z[m]);
                }
            }
            __syncwarp();                                // every lane has read its inputs: the stage is free for the next tile
            if (lane == 0 && more) fetch(tile + tstep);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float4 cv = am_ld4(scb + 16 * m + 4 * t), gv = am_ld4(sg + 16 * m + 4 * t);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float np = slot_np[g + 8 * i];
                    float4 v = am_get4(z[m], i);
                    v.x = am_relu(fmaf(np, gv.x, v.x + cv.x)); v.y = am_relu(fmaf(np, gv.y, v.y + cv.y));
                    v.z = am_relu(fmaf(np, gv.z, v.z + cv.z)); v.w = am_relu(fmaf(np, gv.w, v.w + cv.w));
                    am_put4(z[m], i, v);
                }
            }
            am_to_a<2>(z, ah, al);
            // ---- second layer: l2 = relu(U2^T l1 + c2), one column pair at a time
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                am_pair<4, true>(ah, al, sm + L::B2, 0, 4, 2 * m, lane, z[m]);
                const float4 cv = am_ld4(sc2 + 16 * m + 4 * t);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 v = am_get4(z[m], i);
                    v.x = am_relu(v.x + cv.x); v.y = am_relu(v.y + cv.y); v.z = am_relu(v.z + cv.z); v.w = am_relu(v.w + cv.w);
                    if (rns[i] > 0) am_st4(a.l2 + (int64_t)(base + g + 8 * i) * HID + 16 * m + 4 * t, v);
                    am_put4(z[m], i, v);
                }
            }
            am_to_a<2>(z, ah, al);
            // ---- last message-passing step only: the hidden state itself, h = U3^T l2 + c3
            if (write_h) {
                for (int m = 0; m < 3; ++m) {
                    float ha[2][2][4];
                    am_pair<4, true>(ah, al, sm + L::BH, 0, 6, 2 * m, lane, ha);
                    const float4 cv = am_ld4(sc3 + 16 * m + 4 * t);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 v = am_get4(ha, i);
                        v.x += cv.x; v.y += cv.y; v.z += cv.z; v.w += cv.w;
                        if (rns[i] > 0) am_st4(a.h + (int64_t)(base + g + 8 * i) * HD + 16 * m + 4 * t, v);
                    }
                }
            }
        } else if (proj_gemm) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) {            // l2 of the last message-passing step
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rns[i] > 0) v = am_ld4(st_l2 + (g + 8 * i) * HID + 16 * m + 4 * t);
                    am_put4(z[m], i, v);
                }
            __syncwarp();
            if (lane == 0 && more) fetch(tile + tstep);
            am_to_a<2>(z, ah, al);
        } else {                                         // no product in this launch: nothing was staged
            if (lane == 0 && more) fetch(tile + tstep);
        }

        if (do_proj) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {                // column pairs 0, 1 -> u (a_i block), 2, 3 -> v (a_j block, + b1)
                float acc[2][2][4];
                if (proj_gemm) am_pair<4, true>(ah, al, sm + L::B3, 0, 8, 2 * m, lane, acc);
                else am_zero(acc);
                float* dst = m < 2 ? a.u : a.v;
                const float4 aq = am_ld4(saq + 16 * m + 4 * t);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (rns[i] > 0) {
                        const float4 ax = am_ld4(sAx + slot_sp[g + 8 * i] * 64 + 16 * m + 4 * t);
                        const float qv = slot_q[g + 8 * i];
                        float4 v = am_get4(acc, i);
                        v.x += fmaf(qv, aq.x, ax.x); v.y += fmaf(qv, aq.y, ax.y); v.z += fmaf(qv, aq.z, ax.z); v.w += fmaf(qv, aq.w, ax.w);
                        am_st4(dst + (int64_t)(base + g + 8 * i) * HID + 16 * (m & 1) + 4 * t, v);
                    }
                }
            }
        }
        __syncwarp();
    }
}

#ifndef EPNN_CPU_EMU
template <bool SCOPED, bool UPD>
static cudaError_t launch_atom_mma_t(const Workspace& w, const AtomArgs<float, float>& aa, cudaStream_t st) {
    typedef AmL<UPD> L;
    const size_t smem = sizeof(unsigned) * L::WORDS;
    cudaError_t e = cudaFuncSetAttribute(atom_mma_kernel<SCOPED, UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = div_up(div_up(w.n_atoms, 32), L::NW);
    if (grid > 2 * w.sm_count) grid = 2 * w.sm_count;
    atom_mma_kernel<SCOPED, UPD><<<grid, L::NW * 32, smem, st>>>(aa);
    return cudaGetLastError();
}
cudaError_t launch_atom_mma(const Workspace& w, const AtomArgs<float, float>& aa, cudaStream_t st, int* nl) {
    ++*nl;
    if (aa.mode & ATOM_UPDATE) return aa.scope ? launch_atom_mma_t<true, true>(w, aa, st) : launch_atom_mma_t<false, true>(w, aa, st);
    return aa.scope ? launch_atom_mma_t<true, false>(w, aa, st) : launch_atom_mma_t<false, false>(w, aa, st);
}
#endif
