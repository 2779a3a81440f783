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

__global__ void __launch_bounds__(BP_NW * 32) bundle_count_kernel(const BundlePrepArgs a) {
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

__global__ void __launch_bounds__(BP_NW * 32) bundle_fill_kernel(const BundlePrepArgs a) {
    __shared__ int s_sp[BP_NW][BUNDLE_ATOMS];
    __shared__ unsigned long long s_spm[BP_NW][MAX_SPECIES];
    __shared__ unsigned long long s_mask[BP_NW][BUNDLE_ATOMS];
    __shared__ int s_ust[BP_NW][BUNDLE_ATOMS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* ssp = s_sp[warp];
    unsigned long long* spm = s_spm[warp];
    unsigned long long* smask = s_mask[warp];
    int* sust = s_ust[warp];
    const int nb = a.n_bundles;
    for (int b = blockIdx.x * BP_NW + warp; b < nb; b += gridDim.x * BP_NW) {
        const int2 bd = a.bundle[b];
        const int b0 = bd.x, nat = bd.y;
        for (int f = lane; f < nat; f += 32) { ssp[f] = a.species[b0 + f] & (MAX_SPECIES - 1); smask[f] = a.mask[b0 + f]; }
        __syncwarp();
        bp_species_masks(ssp, nat, lane, spm);
        __syncwarp();
        // ---- the rows' offsets: bundle base (scan over bundles) + warp prefix sum over the rows, 32 rows per round
        int base_nnz = a.boff[b], base_p = a.boff[(nb + 1) + b], base_far = a.boff[2 * (nb + 1) + b], base_far0 = a.boff[3 * (nb + 1) + b];
        int my_rp[2] = {0, 0}, my_us[2] = {0, 0}, my_fo[2] = {0, 0}, my_f0[2] = {0, 0};
#pragma unroll
        for (int rd = 0; rd < (BUNDLE_ATOMS + 31) / 32; ++rd) {
            if (rd * 32 >= nat) break;
            const int r = rd * 32 + lane;
            int deg = 0, dU = 0, far = 0, far0 = 0;
            if (r < nat) {
                const BpRow w = bp_row(a, b0, r);
                const unsigned long long m = smask[r];
                deg = bp_popc(m); dU = bp_popc(m & ~bp_below(r + 1));
                far = (w.a1 - w.a0 - deg) + w.pad;
                far0 = bp_far0_count(spm, ~m & w.sysmask, w.pad);
            }
            int x0 = deg, x1 = dU, x2 = far, x3 = far0;
            for (int o = 1; o < 32; o <<= 1) {
                const int y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
                const int y2 = __shfl_up_sync(0xffffffffu, x2, o), y3 = __shfl_up_sync(0xffffffffu, x3, o);
                if (lane >= o) { x0 += y0; x1 += y1; x2 += y2; x3 += y3; }
            }
            my_rp[rd] = base_nnz + x0 - deg; my_us[rd] = base_p + x1 - dU; my_fo[rd] = base_far + x2 - far; my_f0[rd] = base_far0 + x3 - far0;
            if (r < nat) {
                const int i = b0 + r;
                a.deg[i] = deg; a.degU[i] = dU;
                a.rowptr[i] = my_rp[rd]; a.ustart[i] = my_us[rd]; a.far_off[i] = my_fo[rd]; a.far0_off[i] = my_f0[rd];
                a.atom_b0[i] = b0;
                sust[r] = my_us[rd];
            }
            base_nnz += __shfl_sync(0xffffffffu, x0, 31); base_p += __shfl_sync(0xffffffffu, x1, 31);
            base_far += __shfl_sync(0xffffffffu, x2, 31); base_far0 += __shfl_sync(0xffffffffu, x3, 31);
        }
        if (lane == 0) {
            a.bundle_nat[b0] = nat;
            if (b == nb - 1) {          // closing entries of the offset arrays = the chunk totals
                a.rowptr[a.n_atoms] = a.boff[nb]; a.ustart[a.n_atoms] = a.boff[(nb + 1) + nb];
                a.far_off[a.n_atoms] = a.boff[2 * (nb + 1) + nb]; a.far0_off[a.n_atoms] = a.boff[3 * (nb + 1) + nb];
            }
        }
        __syncwarp();
        // ---- the lists
#pragma unroll
        for (int rd = 0; rd < (BUNDLE_ATOMS + 31) / 32; ++rd) {
            const int r = rd * 32 + lane;
            if (r >= nat) continue;
            const int i = b0 + r;
            const BpRow w = bp_row(a, b0, r);
            const unsigned long long m = smask[r];
            // CSR row (columns ascending) + this row's unordered pairs (its uppers, ascending)
            int k = my_rp[rd], pu = my_us[rd];
            for (unsigned long long rest = m; rest; rest &= rest - 1ull) {
                const int j = bp_ffs(rest) - 1;
                a.col[k] = b0 + j;
                a.rowl[k] = (unsigned char)r;
                if (j > r) {
                    a.pid[k] = pu;
                    a.pair_i[pu] = i; a.pair_j[pu] = b0 + j;      // (its distance: edge_desc_kernel, one thread per pair)
                    ++pu;
                } else {                // the pair is listed under row j: rank of r among j's uppers
                    a.pid[k] = sust[j] + bp_popc(smask[j] & ~bp_below(j + 1) & bp_below(r));
                }
                ++k;
            }
            // far list: the complement of the row inside its system, ascending (the row itself included), then the pad slot
            const int hi = r << 8;
            int wf = my_fo[rd];
            for (unsigned long long rest = ~m & w.sysmask; rest; rest &= rest - 1ull) a.far_list[wf++] = (unsigned short)(hi | (bp_ffs(rest) - 1));
            if (w.pad) a.far_list[wf] = (unsigned short)(hi | 0xFF);
            // species-compressed far list: one slot per species with far columns, species ascending; weight = their number
            int w0 = my_f0[rd];
            const unsigned long long farmask = ~m & w.sysmask;
#pragma unroll
            for (int sp = 0; sp < MAX_SPECIES; ++sp) {
                const int c = bp_popc(spm[sp] & farmask);
                if (c > 0) {
                    a.far0_list[w0] = (unsigned short)(hi | (bp_ffs(spm[sp] & w.sysmask) - 1));
                    a.far0_w[w0] = (unsigned char)c;
                    ++w0;
                }
            }
            if (w.pad) { a.far0_list[w0] = (unsigned short)(hi | 0xFF); a.far0_w[w0] = 0; }
            a.rep[i] = b0 + bp_ffs(spm[w.sp] & w.sysmask) - 1;
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
