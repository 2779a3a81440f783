// Fused list building for chunks that hold only small systems (n <= 48 atoms: the batched-molecule path).
//
// The general path (epnn_neighbor.cu + the far-list kernels of epnn_bundle.cu) builds the lists of a chunk with about
// twenty thread-per-atom kernels and five scans over all atoms; on a million QM9-shaped molecules that was 10 of the
// 17 ms of the neighbour phase, every kernel bound by its own chain of dependent loads.  Here ONE WARP OWNS ONE BUNDLE (whole
// systems, <= 48 atoms): the bundle's coordinates and species are staged once in shared memory with coalesced loads, and
// everything else follows from one 48-bit mask per row ("which atoms of my bundle are e != 0 neighbours"):
//   bundle_count_kernel  distances -> mask[i] (global, 8 B per atom) + the bundle's four totals (CSR entries, unordered
//                        pairs, far slots, species-compressed far slots)
//   4 scans over BUNDLES (not atoms) -> the bundle's base offsets; the last entries are the chunk totals the host needs
//   bundle_fill_kernel   per row: popcounts of the mask -> its four counts, a warp prefix sum -> rowptr / ustart / far_off /
//                        far0_off; then the CSR (col, pid, rowl), the unordered pair list (pair_i, pair_j), the far list,
//                        the species-compressed far list (one popcount per species) and rep -- in one pass.
//   edge_desc_kernel     (epnn_neighbor.cu, one thread per pair) evaluates the pair's float64 distance itself.
// Same definitions, same order, same float64 distance arithmetic as the general path (reference charge_gn.py:122-163,
// :90-94): the lists are bit-identical to it (tests/test_gpu_parity.py neighbour tests, tests/test_emu_bundle_prep.py).
#include "epnn_internal.cuh"

#define BP_NW 8

struct BundlePrepArgs {
    int n_bundles, n_atoms; const int2* bundle;
    const int* atom_sys; const int* sys_off; const int* npad; const int* species; const float* xyz;
    unsigned long long* mask;            // [n_atoms] bit j: atom (first atom of the bundle) + j is a neighbour of the row
    int* btot;                           // count pass:  [4][n_bundles]      nnz | P | far | far0 per bundle
    const int* boff;                     // fill pass:   [4][n_bundles + 1]  their exclusive scans
    int* deg; int* degU; int* rowptr; int* ustart; int* far_off; int* far0_off; int* rep; int* atom_b0; int* bundle_nat;
    int* col; int* pid; unsigned char* rowl; int* pair_i; int* pair_j;
    unsigned short* far_list; unsigned short* far0_list; unsigned char* far0_w;
};

// (defined in epnn_neighbor.cu; repeated here because device functions are per translation unit)
__device__ __forceinline__ double bp_dist64(float xi, float yi, float zi, float xj, float yj, float zj) {
    const double dx = fabs(__dsub_rn((double)xj, (double)xi));
    const double dy = fabs(__dsub_rn((double)yj, (double)yi));
    const double dz = fabs(__dsub_rn((double)zj, (double)zi));
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return __dsqrt_rn(s);
}
__device__ __forceinline__ int bp_popc(unsigned long long m) { return __popc((unsigned)m) + __popc((unsigned)(m >> 32)); }
__device__ __forceinline__ int bp_ffs(unsigned long long m) { const unsigned lo = (unsigned)m; return lo ? __ffs(lo) : 32 + __ffs((unsigned)(m >> 32)); }   // 1-based
__device__ __forceinline__ unsigned long long bp_below(int k) { return k >= 64 ? ~0ull : (1ull << k) - 1ull; }     // bits [0, k)

// Per-row quantities shared by both passes.  r = row inside the bundle, [a0, a1) = its system inside the bundle.
struct BpRow { int a0, a1, sp, pad; unsigned long long sysmask; };
__device__ __forceinline__ BpRow bp_row(const BundlePrepArgs& a, int b0, int r) {
    BpRow w;
    const int s = a.atom_sys[b0 + r];
    w.a0 = a.sys_off[s] - b0; w.a1 = a.sys_off[s + 1] - b0;
    w.sp = a.species[b0 + r] & (MAX_SPECIES - 1);
    w.pad = a.npad[s] > w.a1 - w.a0 ? 1 : 0;
    w.sysmask = bp_below(w.a1) & ~bp_below(w.a0);
    return w;
}
// slots of the species-compressed far list of a row: species with at least one far column (the row itself counts as far) + pad
__device__ __forceinline__ int bp_far0_count(const unsigned long long* spm, unsigned long long farmask, int pad) {
    int n = pad;
#pragma unroll
    for (int k = 0; k < MAX_SPECIES; ++k) n += (spm[k] & farmask) != 0ull;
    return n;
}
// spm[k] = atoms of the bundle with species k (bit = local index); lanes 0..15 build one mask each from the staged species
__device__ __forceinline__ void bp_species_masks(const int* ssp, int nat, int lane, unsigned long long* spm) {
    if (lane < MAX_SPECIES) {
        unsigned long long m = 0ull;
        for (int j = 0; j < nat; ++j) m |= (unsigned long long)(ssp[j] == lane) << j;
        spm[lane] = m;
    }
}

__global__ void __launch_bounds__(BP_NW * 32, 3) bundle_count_kernel(const BundlePrepArgs a) {
    __shared__ float s_xyz[BP_NW][3 * BUNDLE_ATOMS];
    __shared__ int s_sp[BP_NW][BUNDLE_ATOMS];
    __shared__ unsigned long long s_spm[BP_NW][MAX_SPECIES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sx = s_xyz[warp];
    int* ssp = s_sp[warp];
    unsigned long long* spm = s_spm[warp];
    for (int b = blockIdx.x * BP_NW + warp; b < a.n_bundles; b += gridDim.x * BP_NW) {
        const int2 bd = a.bundle[b];
        const int b0 = bd.x, nat = bd.y;
        for (int f = lane; f < 3 * nat; f += 32) sx[f] = a.xyz[3 * (int64_t)b0 + f];
        for (int f = lane; f < nat; f += 32) ssp[f] = a.species[b0 + f] & (MAX_SPECIES - 1);
        __syncwarp();
        bp_species_masks(ssp, nat, lane, spm);
        __syncwarp();
        int t_nnz = 0, t_p = 0, t_far = 0, t_far0 = 0;
        for (int r = lane; r < nat; r += 32) {
            const BpRow w = bp_row(a, b0, r);
            const float xi = sx[3 * r], yi = sx[3 * r + 1], zi = sx[3 * r + 2];
            unsigned long long m = 0ull;
            for (int j = w.a0; j < w.a1; ++j) {
                if (j == r) continue;
                // float32 squared distance (relative error < 1e-6) decides all but a thin shell around the cutoff; only there
                // the reference's own float64 arithmetic (scipy distance_matrix, charge_gn.py:124) is evaluated
                const float xj = sx[3 * j], yj = sx[3 * j + 1], zj = sx[3 * j + 2];
                const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
                const float d2 = dx * dx + dy * dy + dz * dz;
                if (d2 > 9.001f) continue;
                if (d2 < 8.999f || bp_dist64(xi, yi, zi, xj, yj, zj) < 3.0) m |= 1ull << j;
            }
            a.mask[b0 + r] = m;
            const int deg = bp_popc(m);
            t_nnz += deg;
            t_p += bp_popc(m & ~bp_below(r + 1));
            t_far += (w.a1 - w.a0 - deg) + w.pad;
            t_far0 += bp_far0_count(spm, ~m & w.sysmask, w.pad);
        }
        for (int o = 16; o > 0; o >>= 1) {
            t_nnz += __shfl_xor_sync(0xffffffffu, t_nnz, o); t_p += __shfl_xor_sync(0xffffffffu, t_p, o);
            t_far += __shfl_xor_sync(0xffffffffu, t_far, o); t_far0 += __shfl_xor_sync(0xffffffffu, t_far0, o);
        }
        if (lane == 0) {
            a.btot[b] = t_nnz; a.btot[a.n_bundles + b] = t_p; a.btot[2 * a.n_bundles + b] = t_far; a.btot[3 * a.n_bundles + b] = t_far0;
        }
        __syncwarp();
    }
}

// k-th (0-based) set bit of a 64-bit mask (k < popcount): halving steps on popcounts
__device__ __forceinline__ int bp_select(unsigned long long m, int k) {
    unsigned w = (unsigned)m;
    int pos = 0;
    const int c = __popc(w);
    if (k >= c) { k -= c; w = (unsigned)(m >> 32); pos = 32; }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const unsigned part = w & ((1u << s) - 1u);
        const int cc = __popc(part);
        if (k >= cc) { k -= cc; w >>= s; pos += s; } else w = part;
    }
    return pos;
}
// row that owns list entry e: the last r in [0, nat) with off[r] <= e (off[nat] = the bundle's total > e)
__device__ __forceinline__ int bp_find(const int* off, int nat, int e) {
    int lo = 0, hi = nat;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= e) lo = mid; else hi = mid; }
    return lo;
}

// Fill pass.  Phase 1 (lane = row): the rows' four counts and offsets -> the per-atom arrays (coalesced) and shared memory.
// Phase 2 (lane = LIST ENTRY): every list of the bundle is one contiguous range of global memory, so entry e of a list is
// written by lane e mod 32 -- whole 128-byte lines per warp store instead of one 4-byte store per row and instruction (the
// first version: 200 us per 50 k molecules, bound by partial-sector writes); the entry finds its row by bisection of the
// row offsets and its column as the k-th set bit of the row's mask.
__global__ void __launch_bounds__(BP_NW * 32) bundle_fill_kernel(const BundlePrepArgs a) {
    __shared__ int s_sp[BP_NW][BUNDLE_ATOMS];
    __shared__ unsigned long long s_spm[BP_NW][MAX_SPECIES];
    __shared__ unsigned long long s_mask[BP_NW][BUNDLE_ATOMS], s_farm[BP_NW][BUNDLE_ATOMS], s_sysm[BP_NW][BUNDLE_ATOMS];
    __shared__ int s_off[BP_NW][4][BUNDLE_ATOMS + 1];          // local offsets of the rows: CSR | pairs | far | far0; [nat] = total
    __shared__ unsigned short s_spres[BP_NW][BUNDLE_ATOMS];    // species with far columns (bit = species)
    __shared__ unsigned char s_pad[BP_NW][BUNDLE_ATOMS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* ssp = s_sp[warp];
    unsigned long long* spm = s_spm[warp];
    unsigned long long* smask = s_mask[warp];
    unsigned long long* sfar = s_farm[warp];
    unsigned long long* ssys = s_sysm[warp];
    int* o_rp = s_off[warp][0]; int* o_us = s_off[warp][1]; int* o_fo = s_off[warp][2]; int* o_f0 = s_off[warp][3];
    unsigned short* spres = s_spres[warp];
    unsigned char* spad = s_pad[warp];
    const int nb = a.n_bundles;
    for (int b = blockIdx.x * BP_NW + warp; b < nb; b += gridDim.x * BP_NW) {
        const int2 bd = a.bundle[b];
        const int b0 = bd.x, nat = bd.y;
        for (int f = lane; f < nat; f += 32) { ssp[f] = a.species[b0 + f] & (MAX_SPECIES - 1); smask[f] = a.mask[b0 + f]; }
        __syncwarp();
        bp_species_masks(ssp, nat, lane, spm);
        __syncwarp();
        // ---- phase 1: bundle base (scan over bundles) + warp prefix sum over the rows, 32 rows per round
        const int base_nnz = a.boff[b], base_p = a.boff[(nb + 1) + b], base_far = a.boff[2 * (nb + 1) + b], base_far0 = a.boff[3 * (nb + 1) + b];
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;                   // running totals of the rounds done
#pragma unroll
        for (int rd = 0; rd < (BUNDLE_ATOMS + 31) / 32; ++rd) {
            if (rd * 32 >= nat) break;
            const int r = rd * 32 + lane;
            int deg = 0, dU = 0, far = 0, far0 = 0, rep = 0;
            if (r < nat) {
                const BpRow w = bp_row(a, b0, r);
                const unsigned long long m = smask[r], fm = ~m & w.sysmask;
                deg = bp_popc(m); dU = bp_popc(m & ~bp_below(r + 1));
                far = (w.a1 - w.a0 - deg) + w.pad;
                unsigned pres = 0u;
#pragma unroll
                for (int k = 0; k < MAX_SPECIES; ++k) pres |= (unsigned)((spm[k] & fm) != 0ull) << k;
                far0 = __popc(pres) + w.pad;
                sfar[r] = fm; ssys[r] = w.sysmask; spres[r] = (unsigned short)pres; spad[r] = (unsigned char)w.pad;
                rep = b0 + bp_ffs(spm[w.sp] & w.sysmask) - 1;
            }
            int x0 = deg, x1 = dU, x2 = far, x3 = far0;
            for (int o = 1; o < 32; o <<= 1) {
                const int y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
                const int y2 = __shfl_up_sync(0xffffffffu, x2, o), y3 = __shfl_up_sync(0xffffffffu, x3, o);
                if (lane >= o) { x0 += y0; x1 += y1; x2 += y2; x3 += y3; }
            }
            if (r < nat) {
                const int i = b0 + r;
                o_rp[r] = c0 + x0 - deg; o_us[r] = c1 + x1 - dU; o_fo[r] = c2 + x2 - far; o_f0[r] = c3 + x3 - far0;
                a.deg[i] = deg; a.degU[i] = dU;
                a.rowptr[i] = base_nnz + o_rp[r]; a.ustart[i] = base_p + o_us[r]; a.far_off[i] = base_far + o_fo[r]; a.far0_off[i] = base_far0 + o_f0[r];
                a.atom_b0[i] = b0; a.rep[i] = rep;
            }
            c0 += __shfl_sync(0xffffffffu, x0, 31); c1 += __shfl_sync(0xffffffffu, x1, 31);
            c2 += __shfl_sync(0xffffffffu, x2, 31); c3 += __shfl_sync(0xffffffffu, x3, 31);
        }
        if (lane == 0) {
            o_rp[nat] = c0; o_us[nat] = c1; o_fo[nat] = c2; o_f0[nat] = c3;
            a.bundle_nat[b0] = nat;
            if (b == nb - 1) {          // closing entries of the offset arrays = the chunk totals
                a.rowptr[a.n_atoms] = a.boff[nb]; a.ustart[a.n_atoms] = a.boff[(nb + 1) + nb];
                a.far_off[a.n_atoms] = a.boff[2 * (nb + 1) + nb]; a.far0_off[a.n_atoms] = a.boff[3 * (nb + 1) + nb];
            }
        }
        __syncwarp();
        // ---- phase 2: the lists, lane = entry
        // CSR (columns ascending inside a row) with the pair id and the local row of every entry
        for (int e = lane; e < c0; e += 32) {
            const int r = bp_find(o_rp, nat, e), k = e - o_rp[r];
            const unsigned long long m = smask[r];
            const int j = bp_select(m, k);
            a.col[base_nnz + e] = b0 + j;
            a.rowl[base_nnz + e] = (unsigned char)r;
            // j > r: the pair is among the row's own uppers (the tail of the row); else it is listed under row j
            a.pid[base_nnz + e] = base_p + (j > r ? o_us[r] + (k - bp_popc(m & bp_below(r)))
                                                  : o_us[j] + bp_popc(smask[j] & ~bp_below(j + 1) & bp_below(r)));
        }
        // unordered pairs: the uppers of every row, ascending (their distances: edge_desc_kernel, one thread per pair)
        for (int e = lane; e < c1; e += 32) {
            const int r = bp_find(o_us, nat, e);
            a.pair_i[base_p + e] = b0 + r;
            a.pair_j[base_p + e] = b0 + bp_select(smask[r] & ~bp_below(r + 1), e - o_us[r]);
        }
        // far list: the complement of the row inside its system, ascending (the row itself included), then the pad slot
        for (int e = lane; e < c2; e += 32) {
            const int r = bp_find(o_fo, nat, e), k = e - o_fo[r];
            const unsigned long long fm = sfar[r];
            a.far_list[base_far + e] = (unsigned short)((r << 8) | (k < bp_popc(fm) ? bp_select(fm, k) : 0xFF));
        }
        // species-compressed far list: one slot per species with far columns, species ascending; weight = their number
        for (int e = lane; e < c3; e += 32) {
            const int r = bp_find(o_f0, nat, e), k = e - o_f0[r];
            const unsigned pres = spres[r];
            int code = 0xFF, wgt = 0;
            if (k < __popc(pres)) {
                const int sp = bp_select((unsigned long long)pres, k);
                code = bp_ffs(spm[sp] & ssys[r]) - 1;
                wgt = bp_popc(spm[sp] & sfar[r]);
            }
            a.far0_list[base_far0 + e] = (unsigned short)((r << 8) | code);
            a.far0_w[base_far0 + e] = (unsigned char)wgt;
        }
        __syncwarp();
    }
}

// flags[1..4] = nnz, P, n_far, 0 (no large systems in such a chunk) -- what collect_totals_kernel reports on the general path
__global__ void bundle_totals_kernel(const int* __restrict__ boff, int nb, int* __restrict__ flags) {
    flags[1] = boff[nb]; flags[2] = boff[(nb + 1) + nb]; flags[3] = boff[2 * (nb + 1) + nb]; flags[4] = 0;
}

#ifndef EPNN_CPU_EMU
static BundlePrepArgs bp_args(const Workspace& w, const BundlePrepWork& bw) {
    BundlePrepArgs a;
    memset(&a, 0, sizeof(a));
    a.n_bundles = w.n_bundles; a.n_atoms = w.n_atoms; a.bundle = w.bundle;
    a.atom_sys = w.atom_sys; a.sys_off = w.sys_off; a.npad = w.npad; a.species = w.species; a.xyz = w.xyz;
    a.mask = bw.mask; a.btot = bw.btot; a.boff = bw.boff;
    a.deg = w.deg; a.degU = w.degU; a.rowptr = w.rowptr; a.ustart = w.ustart; a.far_off = w.far_off; a.far0_off = w.far0_off;
    a.rep = w.rep; a.atom_b0 = bw.atom_b0; a.bundle_nat = w.bundle_nat;
    a.col = w.col; a.pid = w.pid; a.rowl = w.rowl; a.pair_i = w.pair_i; a.pair_j = w.pair_j;
    a.far_list = w.far_list; a.far0_list = w.far0_list; a.far0_w = w.far0_w;
    return a;
}
static int bp_grid(const Workspace& w) {
    const int g = div_up(w.n_bundles, BP_NW);
    return g < w.sm_count * 8 ? g : w.sm_count * 8;
}

// count pass + the four scans over bundles + the totals into flags[1..4] (the caller copies the flags to the host)
cudaError_t launch_bundle_prep_count(const Workspace& w, const BundlePrepWork& bw, int* scantmp, int* flags, cudaStream_t st, int* nl) {
    const BundlePrepArgs a = bp_args(w, bw);
    bundle_count_kernel<<<bp_grid(w), BP_NW * 32, 0, st>>>(a);
    ++*nl;
    for (int k = 0; k < 4; ++k) {
        cudaError_t e = launch_scan_i32(bw.btot + (size_t)k * w.n_bundles, bw.boff + (size_t)k * (w.n_bundles + 1), w.n_bundles, scantmp, st, nl);
        if (e != cudaSuccess) return e;
    }
    bundle_totals_kernel<<<1, 1, 0, st>>>(bw.boff, w.n_bundles, flags);
    ++*nl;
    return cudaGetLastError();
}

cudaError_t launch_bundle_prep_fill(const Workspace& w, const BundlePrepWork& bw, cudaStream_t st, int* nl) {
    const BundlePrepArgs a = bp_args(w, bw);
    bundle_fill_kernel<<<bp_grid(w), BP_NW * 32, 0, st>>>(a);
    ++*nl;
    return cudaGetLastError();
}
#endif
