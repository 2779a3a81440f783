// Far (e == 0) part of the big-system message sum on the tcgen05 tensor cores, WARP-SPECIALISED -- EXPERIMENTAL
// ("gnn_far_tensor_impl" = 2; the default stays the round-1 kernel, epnn_gnn_tc.cu).  Status at the end of round 2: correct
// (tests/test_gpu_tensor_far.py runs both implementations through the same checks; bit-identical charges to the round-1
// kernel on the 40 k-atom system) but SLOWER: 115 ms against 82 ms for the two live steps of a 40 k-atom system
// (profiles/r02/call27_tc2_ab.log; one group per role: 156 ms; mbarrier.test_wait polling instead of try_wait: 124 ms).
// ncu (call28): 68 % of the stall samples are long-scoreboard waits at the three mbarrier loops -- producers, issuer and
// epilogue wait on each other; with two stages per group the produce -> MMA -> read-back chain of a stage (twelve DEPENDENT
// accumulating MMAs of N = 32 in the middle) is the period, not the instruction count.  Deeper staging needs narrower stages
// (e.g. BF16x3 operands) or N = 64 tiles (two rows per MMA); left as the first item of DESIGN.md section 7.
//
// Same mathematics, operands and 3xTF32 split as gnn_far_tc_kernel (epnn_gnn_tc.cu; reference charge_gn.py:66-70, the unmasked
// reduce_sum over all columns): for a row i and the far columns j,  S_i += relu(W2^T relu(u_i + v_j) + b2), with the 32 x 32
// product of 128 pairs at a time as  D = z_hi W_hi + z_lo W_hi + z_hi W_lo  (twelve tcgen05.mma.kind::tf32, M128 N32 K8, A from
// tensor memory, B = W2^T hi / lo in SWIZZLE_128B shared memory, FP32 accumulation in tensor memory).
//
// Round 1's kernel let the same four warps build A, issue the MMAs (thread 0, after a __syncthreads per row) and run the
// epilogue: ncu showed the tensor pipe 19 % active, the CTA waiting on its own barrier.  Here the three jobs are three sets
// of warps of one persistent CTA per SM, decoupled by mbarriers over TC2_STAGES (4) tensor-memory stages of 96 columns
// (D 32 | A_hi 32 | A_lo 32):
//   warps 0-7   PRODUCERS (two groups of four warps; group g serves rows g and g + 2 of the unit's four)
//                         thread t <-> pair (row r, column j0 + t): z = relu(u_r + v_{j0+t}), split, tcgen05.st into its
//                         TMEM lane; v_{j0+t} stays in registers for the tile, the next tile's row is in flight;
//                         wait empty[s] -> store -> arrive full[s]
//               ISSUER    lane 0 of each group's first warp, after its own arrive: wait full[s] -> 12 x tcgen05.mma ->
//                         tcgen05.commit -> done[s]
//   warps 8-15  EPILOGUE  (two groups likewise) thread t <-> TMEM lane t: wait done[s] -> tcgen05.ld -> arrive empty[s] ->
//                         relu(. + b2), drop the pairs that are near neighbours / beyond the range, add into its own 2 x 32
//                         row sums; per unit a transposing shuffle reduction + a fixed-order sum over the group's four warps
//                         -> S (deterministic)
// A thread cannot overlap two tensor-memory accesses (tcgen05.wait::ld / ::st wait for all of them) and one round trip costs
// ~500 cycles (measured: a single group per role ran at the round-1 kernel's per-CTA rate), hence two groups per role.
// All roles walk the same (unit, tile, row) sequence, so stage and wait parity follow from counters, not from messages.
#include "epnn_internal.cuh"

#define TC2_STAGES 4                    // two per row-parity group (384 of the 512 TMEM columns)
#define TC2_TILE 128
#define TC2_ROWS 4
#define TC2_THREADS 512                 // 2 x 4 producer warps (lane 0 of each group's first warp also issues that group's MMAs), 2 x 4 epilogue warps
#define TC2_SMEM_BYTES (120 * 1024)     // far more than needed (8 KB of B + < 4 KB of the rest): keeps ONE CTA per SM (it allocates all 512 TMEM columns)

struct GnnTc2Args {
    const int* rg_atom; int unit_begin, unit_end, nsplit, n_atoms;
    const int* atom_sys; const int* sys_off;
    const int* rowptr; const int* col;
    const float* u; const float* v;
    const float* Whi; const float* Wlo;        // [32 n][32 k] = hi / lo parts of W2^T, plain row-major
    const float* b2;
    float* S;
    const int* rgl_off; const int* sp_stamp; int stamp;      // far-column de-duplication (epnn_gnn.cu); stamp == 0: off
};

namespace {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start >> 4 | LBO = 1 | SBO = 1024 B >> 4 | version 1 (sm_100) | layout type 2
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                    "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                    "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
// both halves of a stage's D with ONE wait (a thread's TMEM loads all complete together anyway)
__device__ __forceinline__ void ld16x2(uint32_t taddr, float (&d0)[16], float (&d1)[16]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                 " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int c = 0; c < 16; ++c) { d0[c] = __uint_as_float(r[c]); d1[c] = __uint_as_float(r[16 + c]); }
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(bar) : "memory");
}
}   // namespace

__global__ void __launch_bounds__(TC2_THREADS, 1) gnn_far_tc2_kernel(const GnnTc2Args a) {
    extern __shared__ unsigned char tc2_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc2_raw) + 1023) & ~(uintptr_t)1023);
    float* sB = reinterpret_cast<float*>(base);                   // [hi | lo], 1024 floats each, SWIZZLE_128B K-major
    float* sb2 = sB + 2048;                                       // [32]
    float* sRed = sb2 + HID;                                      // [4 warps][4 rows][32]
    uint64_t* sFull = reinterpret_cast<uint64_t*>(sRed + 4 * TC2_ROWS * HID);   // [STAGES] A of the stage is in tensor memory
    uint64_t* sDone = sFull + TC2_STAGES;                         // [STAGES] the stage's MMAs have completed
    uint64_t* sEmpty = sDone + TC2_STAGES;                        // [STAGES] the stage's D has been read back
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sEmpty + TC2_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((HID >> 3) << 17) | ((TC2_TILE >> 4) << 24);

    // ---- one-time setup: B operand, barriers, tensor memory
    for (int f = tid; f < 2 * HID * HID; f += TC2_THREADS) {
        const int part = f >> 10, n = (f >> 5) & 31, k = f & 31;
        const float w = (part ? a.Wlo : a.Whi)[n * HID + k];
        sB[part * 1024 + n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3))] = w;
    }
    if (tid < HID) sb2[tid] = a.b2[tid];
    if (tid == 0) {
        for (int s = 0; s < TC2_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(s32(&sFull[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&sDone[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(s32(&sEmpty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(sTmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *sTmem;

    // the column range of a unit: identical in every role
    auto unit_range = [&](int unit, int& i0, int& a1, int& jlo, int& jhi, int& split) {
        const int rg = unit / a.nsplit;
        split = unit - rg * a.nsplit;
        i0 = a.rg_atom[rg];
        const int sys = a.atom_sys[i0];
        const int a0 = a.sys_off[sys];
        a1 = a.sys_off[sys + 1];
        const int n = a1 - a0;
        int clen = (n + a.nsplit - 1) / a.nsplit;
        clen = (clen + TC2_TILE - 1) / TC2_TILE * TC2_TILE;
        jlo = min(a1, a0 + split * clen);
        jhi = min(a1, jlo + clen);
        if (a.stamp) {                 // the SIMT kernel sums this system's far columns species by species: nothing to do here
            const int ti = a.rgl_off[sys] >> 3;         // (the zero sums are still written: the planes are added per atom)
            if (a.sp_stamp[2 * ti] != a.stamp && a.sp_stamp[2 * ti + 1] == 0) jhi = jlo;
        }
    };

    // Two producer groups and two epilogue groups split the rows by parity (group g: rows g and g + 2): a thread cannot have
    // two tensor-memory accesses in flight (tcgen05.wait::ld / ::st wait for ALL of them), so the ~500-cycle round trip is
    // hidden by a second group working on the next row, not by pipelining inside a thread.  Each group cycles through its
    // OWN two stages (m = the group's row steps so far: stage = 2 g + m % 2, use = m / 2): a stage shared between the groups
    // would let one group run a whole phase ahead of the other, which an mbarrier parity cannot tell from being on time.
    uint32_t tiles_done = 0;
    const int role = warp >> 3;                                   // 0: producers (warps 0-7), 1: epilogue (warps 8-15)
    const int grp = (warp >> 2) & 1;                              // row parity this warp serves
    const int qw = warp & 3;                                      // TMEM lane quarter
    const int pt = (qw << 5) | lane;                              // pair slot of the tile = TMEM lane
    const uint32_t lane_base = tmem + ((uint32_t)(qw * 32) << 16);
    if (role == 0) {
        // ================================================================ PRODUCERS (+ ISSUER: lane 0 of each group's first warp)
        const uint32_t bBase = s32(sB);
        const uint64_t bhi = desc_k128(bBase), blo = desc_k128(bBase + 4096u);
        for (int unit = a.unit_begin + blockIdx.x; unit < a.unit_end; unit += gridDim.x) {
            int i0, a1, jlo, jhi, split;
            unit_range(unit, i0, a1, jlo, jhi, split);
            float vr[32], vn[32];
            auto load_v = [&](float (&dst)[32], int j0) {
                const float4* src = reinterpret_cast<const float4*>(a.v + (int64_t)(j0 + pt) * HID);
                const bool in = j0 + pt < jhi;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 t4 = in ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dst[c * 4] = t4.x; dst[c * 4 + 1] = t4.y; dst[c * 4 + 2] = t4.z; dst[c * 4 + 3] = t4.w;
                }
            };
            if (jlo < jhi) load_v(vn, jlo);
            for (int j0 = jlo; j0 < jhi; j0 += TC2_TILE, ++tiles_done) {
#pragma unroll
                for (int c = 0; c < 32; ++c) vr[c] = vn[c];
                if (j0 + TC2_TILE < jhi) load_v(vn, j0 + TC2_TILE);          // flies during this tile's rows
#pragma unroll 1
                for (int r = grp; r < TC2_ROWS; r += 2) {
                    const uint32_t m = tiles_done * 2 + (r >> 1);            // this group's row steps so far
                    const uint32_t s = grp * 2 + (m & 1u), use = m >> 1;      // each group cycles through its OWN two stages
                    const float4* urow = reinterpret_cast<const float4*>(a.u + (int64_t)min(i0 + r, a1 - 1) * HID);   // same address in every lane
                    bar_wait(s32(&sEmpty[s]), (use & 1u) ^ 1u);              // the stage's previous D has been read back (first use: free)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = lane_base + s * 96u;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float hi[16], lo[16];
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const float4 uu = __ldg(urow + half * 4 + c4);
                            const float uq[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float z = fmaxf(uq[q] + vr[half * 16 + c4 * 4 + q], 0.f);
                                const float h = __uint_as_float(__float_as_uint(z) & 0xFFFFE000u);
                                hi[c4 * 4 + q] = h; lo[c4 * 4 + q] = z - h;
                            }
                        }
                        st16(st + 32u + half * 16u, hi);
                        st16(st + 64u + half * 16u, lo);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    bar_arrive(s32(&sFull[s]));
                    if (qw == 0) {                                           // ISSUER: once the group's 128 rows of A are in tensor memory
                        if (lane == 0) {
                            bar_wait(s32(&sFull[s]), use & 1u);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t d = tmem + s * 96u, ahi = d + 32u, alo = d + 64u;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, ahi + 8 * ks, bhi + 2 * ks, idesc, ks > 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, alo + 8 * ks, bhi + 2 * ks, idesc, 1u);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, ahi + 8 * ks, blo + 2 * ks, idesc, 1u);
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                         :: "r"(s32(&sDone[s])) : "memory");
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ================================================================ EPILOGUE (two groups of four warps <-> TMEM lanes 0-127)
        float* red = sRed + grp * (4 * 2 * HID);                   // [4 warps][2 rows][32] of this group
        for (int unit = a.unit_begin + blockIdx.x; unit < a.unit_end; unit += gridDim.x) {
            int i0, a1, jlo, jhi, split;
            unit_range(unit, i0, a1, jlo, jhi, split);
            float rsum[2][32];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int c = 0; c < 32; ++c) rsum[rr][c] = 0.f;
            // CSR cursors of the group's rows (uniform over the warp): the next near neighbour at or after the current tile
            int cur[2], end[2], nxt[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = grp + 2 * rr;
                cur[rr] = 0; end[rr] = 0; nxt[rr] = 0x7fffffff;
                if (i0 + r < a1) {
                    cur[rr] = a.rowptr[i0 + r]; end[rr] = a.rowptr[i0 + r + 1];
                    while (cur[rr] < end[rr] && a.col[cur[rr]] < jlo) ++cur[rr];
                    if (cur[rr] < end[rr]) nxt[rr] = a.col[cur[rr]];
                }
            }
            for (int j0 = jlo; j0 < jhi; j0 += TC2_TILE, ++tiles_done) {
                const int j = j0 + pt;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = grp + 2 * rr;
                    // is (row r, my column) a near pair?  (rare: only tiles that hold a CSR neighbour of the row look)
                    bool near = false;
                    if (nxt[rr] < j0 + TC2_TILE) {
                        int p = cur[rr];
                        while (p < end[rr]) {
                            const int cc = a.col[p];
                            if (cc >= j0 + TC2_TILE) break;
                            near = near || cc == j;
                            ++p;
                        }
                        cur[rr] = p;
                        nxt[rr] = p < end[rr] ? a.col[p] : 0x7fffffff;
                    }
                    const bool valid = j < jhi && i0 + r < a1 && !near;
                    const uint32_t m = tiles_done * 2 + rr;
                    const uint32_t s = grp * 2 + (m & 1u), use = m >> 1;
                    bar_wait(s32(&sDone[s]), use & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    float d0[16], d1[16];
                    ld16x2(lane_base + s * 96u, d0, d1);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    bar_arrive(s32(&sEmpty[s]));                               // the stage can be refilled while the sums are formed
                    if (valid) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const float4 b0 = *reinterpret_cast<const float4*>(sb2 + c4 * 4), b1 = *reinterpret_cast<const float4*>(sb2 + 16 + c4 * 4);
                            rsum[rr][c4 * 4 + 0] += fmaxf(d0[c4 * 4 + 0] + b0.x, 0.f); rsum[rr][c4 * 4 + 1] += fmaxf(d0[c4 * 4 + 1] + b0.y, 0.f);
                            rsum[rr][c4 * 4 + 2] += fmaxf(d0[c4 * 4 + 2] + b0.z, 0.f); rsum[rr][c4 * 4 + 3] += fmaxf(d0[c4 * 4 + 3] + b0.w, 0.f);
                            rsum[rr][16 + c4 * 4 + 0] += fmaxf(d1[c4 * 4 + 0] + b1.x, 0.f); rsum[rr][16 + c4 * 4 + 1] += fmaxf(d1[c4 * 4 + 1] + b1.y, 0.f);
                            rsum[rr][16 + c4 * 4 + 2] += fmaxf(d1[c4 * 4 + 2] + b1.z, 0.f); rsum[rr][16 + c4 * 4 + 3] += fmaxf(d1[c4 * 4 + 3] + b1.w, 0.f);
                        }
                    }
                }
            }
            // ---- per row: transpose-reduce over the lanes (lane c ends with column c), then the group's four warps in a fixed order
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
                for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
                    for (int q = 0; q < h; ++q) {
                        const float send = (lane & h) ? rsum[rr][q] : rsum[rr][q + h];
                        const float keep = (lane & h) ? rsum[rr][q + h] : rsum[rr][q];
                        rsum[rr][q] = keep + __shfl_xor_sync(0xffffffffu, send, h);
                    }
                }
                red[(qw * 2 + rr) * HID + lane] = rsum[rr][0];
            }
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (pt < 2 * HID) {
                const int rr = pt >> 5, c = pt & 31, r = grp + 2 * rr;
                if (i0 + r < a1) {
                    const float t = ((red[(0 * 2 + rr) * HID + c] + red[(1 * 2 + rr) * HID + c]) + red[(2 * 2 + rr) * HID + c]) + red[(3 * 2 + rr) * HID + c];
                    a.S[((int64_t)split * a.n_atoms + i0 + r) * HID + c] = t;
                }
            }
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");      // red is free for the next unit
        }
    }
    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

cudaError_t launch_gnn_far_tc2(const Workspace& w, const float* Whi, const float* Wlo, const float* b2, int nsplit_tc,
                               cudaStream_t st, int* nl) {
    if (w.n_rg_large == 0) return cudaSuccess;
    GnnTc2Args ga;
    ga.rg_atom = w.rg_large; ga.nsplit = nsplit_tc; ga.n_atoms = w.n_atoms;
    ga.atom_sys = w.atom_sys; ga.sys_off = w.sys_off; ga.rowptr = w.rowptr; ga.col = w.col;
    ga.u = (const float*)w.u; ga.v = (const float*)w.v; ga.Whi = Whi; ga.Wlo = Wlo; ga.b2 = b2; ga.S = (float*)w.S;
    ga.rgl_off = w.rgl_off; ga.sp_stamp = w.sp_stamp; ga.stamp = w.stamp;
    ga.unit_begin = w.rg_begin * nsplit_tc; ga.unit_end = w.rg_end * nsplit_tc;          // sharded call: the row groups overlapping this rank's slice
    int grid = ga.unit_end - ga.unit_begin;
    if (grid < 1) return cudaSuccess;
    if (grid > w.sm_count) grid = w.sm_count;          // persistent: one CTA per SM (it owns the SM's tensor memory)
    cudaError_t e = cudaFuncSetAttribute(gnn_far_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    gnn_far_tc2_kernel<<<grid, TC2_THREADS, TC2_SMEM_BYTES, st>>>(ga);
    ++*nl;
    return cudaGetLastError();
}
