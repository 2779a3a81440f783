// Tensor-core variant of the FAR (e == 0) part of the big-system message sum -- optional ("gnn_far_tensor" = 1).
//
// For a row i of a big system the reference sums  m_ij = relu(W2^T relu(u_i + v_j) + b2)  over ALL columns j that are
// not within the 3 A cutoff (charge_gn.py:66-70, no mask): O(n^2) pairs, 99.9 % of the work of a 1 M-atom system.
// The inner product is a [pairs x 32] x [32 x 32] GEMM, so it maps onto Blackwell's 5th-generation tensor cores:
//
//   * a CTA (4 warps) takes a row group (4 rows i) and a column range; per tile of 128 consecutive j it stages the
//     v rows in shared memory once and, for each of the 4 rows, thread t builds row t of the A operand
//     z_t = relu(u_i + v_{j0+t}) directly in the UMMA canonical K-major SWIZZLE_128B layout (one 128-byte row per pair);
//   * precision: "3xTF32" error-compensated split  z = z_hi + z_lo  (z_hi = z with the low 13 mantissa bits cleared,
//     z_lo = z - z_hi exactly), W2 likewise on the host:  D = z_hi W_hi + z_lo W_hi + z_hi W_lo, FP32 accumulation in
//     TMEM -- twelve tcgen05.mma (M128 N32 K8, kind::tf32) per 128 pairs, issued by one thread;
//   * two A buffers / two TMEM accumulator stages: the MMAs of one row overlap the epilogue of the previous one.
//     Completion is signalled with tcgen05.commit -> mbarrier; the epilogue reads its pair's 32 outputs with one
//     tcgen05.ld (32x32b.x32: thread t of warp w owns TMEM lane 32 w + t), applies relu(. + b2), drops the pairs that
//     belong to the near list / lie beyond the system, and a 31-shuffle transpose-reduce leaves lane c with column c's
//     sum over the warp's 32 pairs.  Per-row sums are combined across the 4 warps in a fixed order: deterministic.
//
// The near (e != 0) pairs and the pad pseudo-pair stay on the FP32 SIMT kernel (epnn_gnn.cu, skip_far mode), written to
// their own partial-sum plane; the per-atom kernel adds the planes in a fixed order.
#include <type_traits>

#include "epnn_internal.cuh"

#define TC_THREADS 128
#define TC_TILE 128                    // pairs (MMA M) per tile
#define TC_ROWS 4                      // rows i per work unit (= one row group)

struct GnnTcArgs {
    const int* rg_atom; int unit_begin, unit_end, nsplit, n_atoms;
    const int* atom_sys; const int* sys_off;
    const int* rowptr; const int* col;
    const float* u; const float* v;
    const float* Whi; const float* Wlo;        // [32 n][32 k] = hi / lo parts of W2^T, plain row-major
    const float* b2;
    float* S;
    const int* rgl_off; const int* sp_stamp; int stamp;      // far-column de-duplication (epnn_gnn.cu); stamp == 0: off
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start >> 4 | LBO = 1 | SBO = 1024 B >> 4 | version 1 (sm_100) | layout type 2
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// A operand from tensor memory (lanes = pairs, 8 consecutive 32-bit columns per K step), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                    "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                    "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&d)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int c = 0; c < 16; ++c) d[c] = __uint_as_float(r[c]);
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&d)[32]) {
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
    for (int c = 0; c < 32; ++c) d[c] = __uint_as_float(r[c]);
}

// shared-memory carve-up (byte offsets from a 1024-aligned base)
#define TC_OFF_A 0                                 // Ahi[2], Alo[2]: 4 x 16 KB
#define TC_OFF_B (4 * 16384)                       // Bhi, Blo: 2 x 4 KB
#define TC_OFF_V (TC_OFF_B + 2 * 4096)             // v tile: 16 KB
#define TC_OFF_MISC (TC_OFF_V + 16384)             // u rows, b2, masks, reduction buffer, barriers, tmem address
#define TC_SMEM_BYTES (TC_OFF_MISC + 4096 + 1024)  // + alignment slack

__global__ void __launch_bounds__(TC_THREADS, 2) gnn_far_tc_kernel(const GnnTcArgs a) {
    extern __shared__ unsigned char tc_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~(uintptr_t)1023);
    float* sB = reinterpret_cast<float*>(base + TC_OFF_B);        // [hi | lo], 1024 floats each
    float* sU = reinterpret_cast<float*>(base + TC_OFF_MISC);     // [4][32]
    float* sb2 = sU + TC_ROWS * HID;                              // [32]
    float* sRed = sb2 + HID;                                      // [4 warps][4 rows][32]
    unsigned* sMask = reinterpret_cast<unsigned*>(sRed + 4 * TC_ROWS * HID);   // [2 tile parities][4 rows][4 words]
    int* sPtr = reinterpret_cast<int*>(sMask + 2 * TC_ROWS * 4);  // [4] CSR cursors of the 4 rows
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sPtr + 4);       // [2] (8-byte aligned: offsets above are multiples of 8)
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((HID >> 3) << 17) | ((TC_TILE >> 4) << 24);

    // ---- one-time setup: B operand, barriers, tensor memory
    for (int f = tid; f < 2 * HID * HID; f += TC_THREADS) {
        const int part = f >> 10, n = (f >> 5) & 31, k = f & 31;
        const float w = (part ? a.Wlo : a.Whi)[n * HID + k];
        sB[part * 1024 + n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3))] = w;
    }
    if (tid < HID) sb2[tid] = a.b2[tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sBar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sBar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sTmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *sTmem;
    const uint32_t bBase = smem_u32(sB);
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 TMEM lanes
    uint32_t uses0 = 0, uses1 = 0;            // completed phases of the two barriers (uniform across the CTA)

    for (int unit = a.unit_begin + blockIdx.x; unit < a.unit_end; unit += gridDim.x) {
        const int rg = unit / a.nsplit, split = unit - rg * a.nsplit;
        const int i0 = a.rg_atom[rg];
        const int sys = a.atom_sys[i0];
        const int a0 = a.sys_off[sys], a1 = a.sys_off[sys + 1];
        const int n = a1 - a0;
        int clen = (n + a.nsplit - 1) / a.nsplit;
        clen = (clen + TC_TILE - 1) / TC_TILE * TC_TILE;
        const int jlo = min(a1, a0 + split * clen);
        int jhi = min(a1, jlo + clen);
        if (a.stamp) {                 // the SIMT kernel sums this system's far columns species by species: nothing to do here
            const int ti = a.rgl_off[sys] >> 3;         // (the zero sums are still written: the planes are added per atom)
            if (a.sp_stamp[2 * ti] != a.stamp && a.sp_stamp[2 * ti + 1] == 0) jhi = jlo;
        }
        __syncthreads();                                          // previous unit's reduction buffer / u rows are free
        if (tid < TC_ROWS * HID) {
            const int r = tid >> 5;
            sU[tid] = i0 + r < a1 ? a.u[(int64_t)(i0 + r) * HID + (tid & 31)] : 0.f;
        }
        if (tid < TC_ROWS) {                                      // CSR cursor of row r: first column >= jlo
            int p = 0;
            if (i0 + tid < a1) { p = a.rowptr[i0 + tid]; const int e = a.rowptr[i0 + tid + 1]; while (p < e && a.col[p] < jlo) ++p; }
            sPtr[tid] = p;
        }
        __syncthreads();                                          // u rows and CSR cursors visible to every warp
        float rsum[TC_ROWS][32];                                  // this thread's pairs only: row r, column c (registers)
#pragma unroll
        for (int r = 0; r < TC_ROWS; ++r)
#pragma unroll
            for (int c = 0; c < 32; ++c) rsum[r][c] = 0.f;
        bool pending = false;
        int pend_j0 = 0, pend_par = 0;
        // next CSR neighbour column of row (i0 + warp) not yet consumed by a tile (INT_MAX: none left): lets almost
        // every tile skip the mask build without touching global memory
        int row_end = 0, cur_p = 0, next_col = 0x7fffffff;
        if (i0 + warp < a1) {
            row_end = a.rowptr[i0 + warp + 1];
            cur_p = sPtr[warp];
            if (cur_p < row_end) next_col = a.col[cur_p];
        }
        float vr[32];                                             // this thread's v row of the current tile
        auto load_v = [&](int j0) {
            const float4* src = reinterpret_cast<const float4*>(a.v + (int64_t)(j0 + tid) * HID);
            const bool in = j0 + tid < jhi;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 t4 = in ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                vr[c * 4] = t4.x; vr[c * 4 + 1] = t4.y; vr[c * 4 + 2] = t4.z; vr[c * 4 + 3] = t4.w;
            }
        };
        if (jlo < jhi) load_v(jlo);

        // epilogue of (row r, stage r & 1): r is a compile-time constant at every call site (the row loop is unrolled)
        auto epilogue = [&](auto rc, int j0, int par) {
            constexpr int r = decltype(rc)::value;
            constexpr int stage = r & 1;
            mbar_wait(smem_u32(&sBar[stage]), (stage ? uses1 : uses0) & 1u);
            if (stage) ++uses1; else ++uses0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int j = j0 + tid;
            const bool valid = j < jhi && i0 + r < a1 && !((sMask[(par * TC_ROWS + r) * 4 + warp] >> lane) & 1u);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float d[16];
                tmem_ld16(lane_base + (uint32_t)stage * 96u + half * 16u, d);
                if (valid) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 bb = *reinterpret_cast<const float4*>(sb2 + half * 16 + c4 * 4);
                        rsum[r][half * 16 + c4 * 4 + 0] += fmaxf(d[c4 * 4 + 0] + bb.x, 0.f); rsum[r][half * 16 + c4 * 4 + 1] += fmaxf(d[c4 * 4 + 1] + bb.y, 0.f);
                        rsum[r][half * 16 + c4 * 4 + 2] += fmaxf(d[c4 * 4 + 2] + bb.z, 0.f); rsum[r][half * 16 + c4 * 4 + 3] += fmaxf(d[c4 * 4 + 3] + bb.w, 0.f);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        };

        int tile_par = 0;
        for (int j0 = jlo; j0 < jhi; j0 += TC_TILE, tile_par ^= 1) {
            // ---- near masks of this tile (warp r <-> row r): bit (j - j0) set for CSR neighbours j in [j0, j0 + 128)
            if (lane < 4) sMask[(tile_par * TC_ROWS + warp) * 4 + lane] = 0u;
            __syncwarp();
            if (next_col < j0 + TC_TILE) {                        // rare: this tile holds near neighbours of the row
                for (;;) {
                    const int c = cur_p + lane < row_end ? a.col[cur_p + lane] : 0x7fffffff;
                    const bool in = c < j0 + TC_TILE;
                    if (in) atomicOr(&sMask[(tile_par * TC_ROWS + warp) * 4 + ((c - j0) >> 5)], 1u << ((c - j0) & 31));
                    const int cnt = __popc(__ballot_sync(0xffffffffu, in));
                    cur_p += cnt;
                    if (cnt < 32) break;
                }
                next_col = cur_p < row_end ? a.col[cur_p] : 0x7fffffff;
                __syncwarp();
            }
            auto row_step = [&](auto rc) {
                constexpr int r = decltype(rc)::value;
                constexpr int stage = r & 1;
                // ---- produce row t of A (hi and lo) for pair (i0 + r, j0 + t): registers -> tensor memory (lane = pair)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float hi[16], lo[16];
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 uu = *reinterpret_cast<const float4*>(sU + r * HID + half * 16 + c4 * 4);
                        const float uq[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float z = fmaxf(uq[q] + vr[half * 16 + c4 * 4 + q], 0.f);
                            const float h = __uint_as_float(__float_as_uint(z) & 0xFFFFE000u);
                            hi[c4 * 4 + q] = h; lo[c4 * 4 + q] = z - h;
                        }
                    }
                    tmem_st16(lane_base + (uint32_t)stage * 96u + 32u + half * 16u, hi);
                    tmem_st16(lane_base + (uint32_t)stage * 96u + 64u + half * 16u, lo);
                }
                if (r == TC_ROWS - 1 && j0 + TC_TILE < jhi) load_v(j0 + TC_TILE);   // vr is dead now: next tile's row flies during the epilogues
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d = tmem + (uint32_t)stage * 96u, ahi = d + 32u, alo = d + 64u;
                    const uint64_t bhi = umma_desc_k128(bBase), blo = umma_desc_k128(bBase + 4096u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_tf32_ts(d, ahi + 8 * ks, bhi + 2 * ks, idesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_tf32_ts(d, alo + 8 * ks, bhi + 2 * ks, idesc, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_tf32_ts(d, ahi + 8 * ks, blo + 2 * ks, idesc, 1u);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 :: "r"(smem_u32(&sBar[stage])) : "memory");
                }
                // the previous row's epilogue overlaps the MMAs just issued (row 0's predecessor is row 3 of the previous tile)
                if (r > 0) epilogue(std::integral_constant<int, (r + 3) & 3>{}, j0, tile_par);
                else if (pending) epilogue(std::integral_constant<int, 3>{}, pend_j0, pend_par);
            };
            row_step(std::integral_constant<int, 0>{});
            row_step(std::integral_constant<int, 1>{});
            row_step(std::integral_constant<int, 2>{});
            row_step(std::integral_constant<int, 3>{});
            pending = true; pend_j0 = j0; pend_par = tile_par;
        }
        if (pending) epilogue(std::integral_constant<int, 3>{}, pend_j0, pend_par);
        // ---- per row: transpose-reduce over the lanes (lane c ends with column c), then the four warps in a fixed order
#pragma unroll
        for (int r = 0; r < TC_ROWS; ++r) {
#pragma unroll
            for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
                for (int q = 0; q < h; ++q) {
                    const float send = (lane & h) ? rsum[r][q] : rsum[r][q + h];
                    const float keep = (lane & h) ? rsum[r][q + h] : rsum[r][q];
                    rsum[r][q] = keep + __shfl_xor_sync(0xffffffffu, send, h);
                }
            }
            sRed[(warp * TC_ROWS + r) * HID + lane] = rsum[r][0];
        }
        __syncthreads();
        if (tid < TC_ROWS * HID) {
            const int r = tid >> 5, c = tid & 31;
            if (i0 + r < a1) {
                const float t = ((sRed[(0 * TC_ROWS + r) * HID + c] + sRed[(1 * TC_ROWS + r) * HID + c]) +
                                 sRed[(2 * TC_ROWS + r) * HID + c]) + sRed[(3 * TC_ROWS + r) * HID + c];
                a.S[((int64_t)split * a.n_atoms + i0 + r) * HID + c] = t;
            }
        }
    }
    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u) : "memory");
}

cudaError_t launch_gnn_far_tc(const Workspace& w, const float* Whi, const float* Wlo, const float* b2, int nsplit_tc,
                              cudaStream_t st, int* nl) {
    if (w.n_rg_large == 0) return cudaSuccess;
    GnnTcArgs ga;
    ga.rg_atom = w.rg_large; ga.nsplit = nsplit_tc; ga.n_atoms = w.n_atoms;
    ga.atom_sys = w.atom_sys; ga.sys_off = w.sys_off; ga.rowptr = w.rowptr; ga.col = w.col;
    ga.u = (const float*)w.u; ga.v = (const float*)w.v; ga.Whi = Whi; ga.Wlo = Wlo; ga.b2 = b2; ga.S = (float*)w.S;
    ga.rgl_off = w.rgl_off; ga.sp_stamp = w.sp_stamp; ga.stamp = w.stamp;
    ga.unit_begin = w.rg_begin * nsplit_tc; ga.unit_end = w.rg_end * nsplit_tc;          // sharded call: the row groups overlapping this rank's slice
    int grid = ga.unit_end - ga.unit_begin;
    if (grid < 1) return cudaSuccess;
    if (grid > 2 * w.sm_count) grid = 2 * w.sm_count;
    cudaError_t e = cudaFuncSetAttribute(gnn_far_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    gnn_far_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ga);
    ++*nl;
    return cudaGetLastError();
}
