// Neighbour list + radial descriptors.
//
// Replaces charge_gn.get_init_edges (reference charge_gn.py:122-163) and the is_near predicate of
// EPN_layer.call (charge_gn.py:90-94).  The reference materialises dense (n,n,48) float64 tensors; here
// the same arithmetic is evaluated only for pairs with D < 3.0 A (everything else is exactly zero):
//
//   D    = sqrt((dx^2 + dy^2) + dz^2)  in float64 from float32 coordinates, no FMA contraction
//          (scipy.spatial.distance_matrix, charge_gn.py:124)
//   C    = (cos(pi*D/3) + 1)/2, C = 0 for D >= 3, C = 1 for D <= 0, C_ii = 0        (:148-152)
//   e_k  = float32(C * exp(-2 (D - mu_k)^2)),  mu = linspace(0.1, 3.0, 48)            (:123,153-161)
//   near = max_k e_k > float32(1e-5)                                                  (:90-94)
//
// Output: CSR of the e != 0 set (rowptr/col ascending/pid) + per unordered pair (i<j): e[48], near flag.
#include "epnn_internal.cuh"

__constant__ double c_mu[ED];

__constant__ double c_B[ED * EDR];       // orthonormal basis of the descriptor family, [k][r]

#ifdef EPNN_CPU_EMU
void emu_set_rbf(const double* mu, const double* B) { memcpy(c_mu, mu, sizeof(c_mu)); memcpy(c_B, B, sizeof(c_B)); }
#else
cudaError_t upload_rbf_centers(const double* mu) { return cudaMemcpyToSymbol(c_mu, mu, sizeof(double) * ED); }
cudaError_t upload_rbf_basis(const double* B) { return cudaMemcpyToSymbol(c_B, B, sizeof(double) * ED * EDR); }
#endif

// float64 distance exactly as scipy computes it; intrinsics forbid FMA contraction.
__device__ __forceinline__ double dist64(float xi, float yi, float zi, float xj, float yj, float zj) {
    const double dx = fabs(__dsub_rn((double)xj, (double)xi));
    const double dy = fabs(__dsub_rn((double)yj, (double)yi));
    const double dz = fabs(__dsub_rn((double)zj, (double)zi));
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return __dsqrt_rn(s);
}

// Cheap float32 reject: squared distance clearly above 9 (relative fp32 error of d2 is < 1e-6).
__device__ __forceinline__ bool far_reject(float xi, float yi, float zi, float xj, float yj, float zj) {
    const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    return dx * dx + dy * dy + dz * dz > 9.001f;
}

__device__ __forceinline__ double cutoff_fn(double D) {
    if (D >= 3.0) return 0.0;
    if (D <= 0.0) return 1.0;
    const double arg = __ddiv_rn(__dmul_rn(3.141592653589793, D), 3.0);
    return __ddiv_rn(__dadd_rn(cos(arg), 1.0), 2.0);
}

__device__ __forceinline__ float rbf_value(double C, double D, int k) {
    const double d = __dsub_rn(D, c_mu[k]);
    const double g = exp(__dmul_rn(-2.0, __dmul_rn(d, d)));
    return __double2float_rn(__dmul_rn(C, g));
}

// ------------------------------------------------------------------------------------------------
// prep: atom -> system map (binary search over the offsets) and initial charges q0 = fl32(Q)/n.
__global__ void prep_kernel(int n_atoms, int n_sys, const int* __restrict__ sys_off, const float* __restrict__ Qsys,
                            int* __restrict__ atom_sys, double* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    int lo = 0, hi = n_sys;                 // find s with off[s] <= i < off[s+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sys_off[mid] <= i) lo = mid; else hi = mid;
    }
    atom_sys[i] = lo;
    const int n = sys_off[lo + 1] - sys_off[lo];
    q[i] = (double)__fdiv_rn(Qsys[lo], (float)n);      // charge_gn.py:337-338
}

#ifndef EPNN_CPU_EMU
cudaError_t launch_prep(const Workspace& w, cudaStream_t st, int* nl) {
    if (w.n_atoms == 0) return cudaSuccess;
    prep_kernel<<<div_up(w.n_atoms, 256), 256, 0, st>>>(w.n_atoms, w.n_sys, w.sys_off, w.Qsys, w.atom_sys, w.q);
    ++*nl;
    return cudaGetLastError();
}
#endif

// ------------------------------------------------------------------------------------------------
// Cell lists for big systems (n > CELL_MIN atoms).  Every such system gets its own uniform grid with cell edge
// h >= 3.01 A (> the 3.0 A cutoff with a margin far above float32 rounding of the cell coordinate), enlarged if the
// bounding box would need more cells than the host-side budget (4 n + 64).  Atoms are binned with integer atomics
// (only the ORDER inside a cell depends on scheduling; rows are sorted afterwards, so the lists do not).
__global__ void cell_setup_kernel(int n_large, const int* __restrict__ large_sys, const int* __restrict__ large_base,
                                  const int* __restrict__ sys_off, const float* __restrict__ xyz, CellGrid* __restrict__ grid) {
    __shared__ float red[6][8];
    const int s = large_sys[blockIdx.x];
    const int a0 = sys_off[s], a1 = sys_off[s + 1];
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x)
        for (int d = 0; d < 3; ++d) { const float v = xyz[3 * i + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if ((threadIdx.x & 31) == 0) { red[d][threadIdx.x >> 5] = lo[d]; red[3 + d][threadIdx.x >> 5] = hi[d]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int d = 0; d < 3; ++d)
            for (int w = 0; w < nw; ++w) { lo[d] = fminf(lo[d], red[d][w]); hi[d] = fmaxf(hi[d], red[3 + d][w]); }
        const long long budget = 4ll * (a1 - a0) + 64;
        float h = 3.01f;
        int nx, ny, nz;
        for (;;) {
            nx = (int)floorf((hi[0] - lo[0]) / h) + 1; ny = (int)floorf((hi[1] - lo[1]) / h) + 1; nz = (int)floorf((hi[2] - lo[2]) / h) + 1;
            if ((long long)nx * ny * nz <= budget) break;
            h *= 1.26f;                              // ~2x fewer cells per round
        }
        CellGrid g;
        g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.inv_h = 1.0f / h;
        g.nx = nx; g.ny = ny; g.nz = nz; g.base = large_base[blockIdx.x];
        grid[s] = g;
    }
}

__device__ __forceinline__ void cell_of(const CellGrid& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = min(g.nx - 1, max(0, (int)floorf((x - g.ox) * g.inv_h)));
    cy = min(g.ny - 1, max(0, (int)floorf((y - g.oy) * g.inv_h)));
    cz = min(g.nz - 1, max(0, (int)floorf((z - g.oz) * g.inv_h)));
}

// pass 0: histogram (cell_cnt) ; pass 1: scatter into cell order (cell_atoms) using cell_start + a cursor
template <int PASS>
__global__ void cell_bin_kernel(int n_atoms, const int* __restrict__ atom_sys, const int* __restrict__ sys_off,
                                const float* __restrict__ xyz, const CellGrid* __restrict__ grid, int* __restrict__ cell_cnt,
                                const int* __restrict__ cell_start, int* __restrict__ cell_atoms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const int s = atom_sys[i];
    if (sys_off[s + 1] - sys_off[s] <= CELL_MIN) return;
    const CellGrid g = grid[s];
    int cx, cy, cz;
    cell_of(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], cx, cy, cz);
    const int c = g.base + (cz * g.ny + cy) * g.nx + cx;
    const int k = atomicAdd(cell_cnt + c, 1);
    if (PASS == 1) cell_atoms[cell_start[c] + k] = i;
}

#ifndef EPNN_CPU_EMU
cudaError_t launch_cell_build(const Workspace& w, const CellWork& cw, int* scantmp, cudaStream_t st, int* nl) {
    if (cw.n_large == 0) return cudaSuccess;
    cell_setup_kernel<<<cw.n_large, 256, 0, st>>>(cw.n_large, cw.large_sys, cw.large_base, w.sys_off, w.xyz, cw.grid);
    cudaError_t e = cudaMemsetAsync(cw.cell_cnt, 0, sizeof(int) * (size_t)cw.n_cells, st);
    if (e != cudaSuccess) return e;
    cell_bin_kernel<0><<<div_up(w.n_atoms, 256), 256, 0, st>>>(w.n_atoms, w.atom_sys, w.sys_off, w.xyz, cw.grid, cw.cell_cnt, nullptr, nullptr);
    *nl += 2;
    e = launch_scan_i32(cw.cell_cnt, cw.cell_start, cw.n_cells, scantmp, st, nl);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(cw.cell_cnt, 0, sizeof(int) * (size_t)cw.n_cells, st);
    if (e != cudaSuccess) return e;
    cell_bin_kernel<1><<<div_up(w.n_atoms, 256), 256, 0, st>>>(w.n_atoms, w.atom_sys, w.sys_off, w.xyz, cw.grid, cw.cell_cnt, cw.cell_start, cw.cell_atoms);
    ++*nl;
    return cudaGetLastError();
}
#endif

// ------------------------------------------------------------------------------------------------
// Neighbour search: one thread per atom row, two passes (count, fill).  Systems with n <= CELL_MIN scan their own
// atoms (columns come out ascending because j is scanned ascending); bigger systems walk the 27 cells around the
// atom and sort the row afterwards, so both paths emit identical, ascending, bit-exact lists.
#define NBR_BLOCK 128
#define NBR_STAGE 1280                      // atoms whose coordinates fit the staging buffer (15 KB)
template <bool FILL>
__global__ void __launch_bounds__(NBR_BLOCK) nbr_kernel(int n_atoms, const int* __restrict__ atom_sys, const int* __restrict__ sys_off,
                           const float* __restrict__ xyz, int* __restrict__ deg, int* __restrict__ degU,
                           const int* __restrict__ rowptr, const int* __restrict__ ustart, int* __restrict__ col,
                           int* __restrict__ pair_i, int* __restrict__ pair_j, double* __restrict__ pair_D,
                           const CellGrid* __restrict__ grid, const int* __restrict__ cell_start,
                           const int* __restrict__ cell_atoms, double* __restrict__ Dtmp, int row_lo, int row_hi) {
    // Coordinates of the block's systems, staged once in shared memory with coalesced loads (the AoS float3 array is read as
    // a flat run of floats): the block's atoms are consecutive, so the systems they belong to cover one contiguous range of
    // atoms -- at most NBR_BLOCK + 2 * 511 of them unless a cell-list system (n > CELL_MIN) is involved, in which case that
    // system's rows read global memory through the cell list as before.
    __shared__ float sxyz[3 * NBR_STAGE];
    const int i_first = blockIdx.x * blockDim.x;
    const int i_last = min(n_atoms, i_first + (int)blockDim.x) - 1;
    const int st_lo = sys_off[atom_sys[i_first]];
    const int st_hi = sys_off[atom_sys[i_last] + 1];
    const bool staged = st_hi - st_lo <= NBR_STAGE;
    if (staged) {
        const float* src = xyz + 3 * (int64_t)st_lo;
        for (int f = threadIdx.x; f < 3 * (st_hi - st_lo); f += blockDim.x) sxyz[f] = src[f];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const int s = atom_sys[i];
    const int a0 = sys_off[s], a1 = sys_off[s + 1];
    // Sharded call (epnn_shard_init): rows of LARGE systems are built by the rank that owns them, [row_lo, row_hi); small
    // systems are replicated.  A pair is listed under row i if j > i, or if j lies below the slice ("foreign lower": row j
    // is not built here) -- always as (min, max), so every rank evaluates a cut pair in the same canonical orientation.
    const bool large = a1 - a0 > SMALL_MAX;
    if (large && (i < row_lo || i >= row_hi)) { if (!FILL) { deg[i] = 0; degU[i] = 0; } return; }
    const int flo = large ? row_lo : a0;            // columns below flo are foreign lowers (none for small systems: j >= a0)
    const float xi = xyz[3 * i], yi = xyz[3 * i + 1], zi = xyz[3 * i + 2];
    int c = 0, cu = 0;
    int wp = 0, wu = 0;
    if (FILL) { wp = rowptr[i]; wu = ustart[i]; }
    if (a1 - a0 <= CELL_MIN) {
        const float* cx = staged ? sxyz - 3 * (int64_t)st_lo : xyz;      // same values either way: the lists cannot differ
        for (int j = a0; j < a1; ++j) {
            if (j == i) continue;
            const float xj = cx[3 * j], yj = cx[3 * j + 1], zj = cx[3 * j + 2];
            if (far_reject(xi, yi, zi, xj, yj, zj)) continue;
            const double D = dist64(xi, yi, zi, xj, yj, zj);
            if (D < 3.0) {
                const bool listed = j > i || j < flo;
                if (FILL) {
                    col[wp + c] = j;
                    if (listed) { pair_i[wu + cu] = min(i, j); pair_j[wu + cu] = max(i, j); pair_D[wu + cu] = D; }
                }
                ++c;
                cu += listed;
            }
        }
    } else {
        const CellGrid g = grid[s];
        int cx, cy, cz;
        cell_of(g, xi, yi, zi, cx, cy, cz);
        for (int dz = -1; dz <= 1; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= g.nz) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= g.ny) continue;
                const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
                const int cb = g.base + (z * g.ny + y) * g.nx;
                const int k0 = cell_start[cb + x0], k1 = cell_start[cb + x1 + 1];     // x-neighbours are contiguous cells
                for (int k = k0; k < k1; ++k) {
                    const int j = cell_atoms[k];
                    if (j == i) continue;
                    const float xj = xyz[3 * j], yj = xyz[3 * j + 1], zj = xyz[3 * j + 2];
                    if (far_reject(xi, yi, zi, xj, yj, zj)) continue;
                    const double D = dist64(xi, yi, zi, xj, yj, zj);
                    if (D < 3.0) {
                        if (FILL) { col[wp + c] = j; Dtmp[wp + c] = D; }
                        ++c;
                        cu += (j > i || j < flo);
                    }
                }
            }
        }
        if (FILL) {
            for (int p = 1; p < c; ++p) {                       // insertion sort of the row by column index
                const int jv = col[wp + p];
                const double dv = Dtmp[wp + p];
                int q = p - 1;
                while (q >= 0 && col[wp + q] > jv) { col[wp + q + 1] = col[wp + q]; Dtmp[wp + q + 1] = Dtmp[wp + q]; --q; }
                col[wp + q + 1] = jv; Dtmp[wp + q + 1] = dv;
            }
            int k = 0;
            for (int p = 0; p < c; ++p) {
                const int j = col[wp + p];
                if (j > i || j < flo) { pair_i[wu + k] = min(i, j); pair_j[wu + k] = max(i, j); pair_D[wu + k] = Dtmp[wp + p]; ++k; }
            }
        }
    }
    if (!FILL) { deg[i] = c; degU[i] = cu; }
}

#ifndef EPNN_CPU_EMU
cudaError_t launch_nbr_count(const Workspace& w, const CellWork& cw, cudaStream_t st, int* nl) {
    if (w.n_atoms == 0) return cudaSuccess;
    nbr_kernel<false><<<div_up(w.n_atoms, NBR_BLOCK), NBR_BLOCK, 0, st>>>(w.n_atoms, w.atom_sys, w.sys_off, w.xyz, w.deg, w.degU,
                                                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                              cw.grid, cw.cell_start, cw.cell_atoms, nullptr, w.row_lo, w.row_hi);
    ++*nl;
    return cudaGetLastError();
}
#endif

// pid of every CSR entry.  Row r lists degU[r] pairs in column order: its foreign lowers (sharded calls only: columns
// below the slice, always at the head of the ascending row) and its uppers (col > r, the tail of the row).  An entry
// (i, j) of row i is either listed there (rank among the listed entries of row i), or j is an owned lower: then the pair
// sits among the uppers of row j (binary search; rows are sorted).
__global__ void nbr_rev_kernel(int n_atoms, const int* __restrict__ rowptr, const int* __restrict__ ustart,
                               const int* __restrict__ degU, const int* __restrict__ col, int* __restrict__ pid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const int r0 = rowptr[i], r1 = rowptr[i + 1];
    int nup = 0;
    for (int k = r1 - 1; k >= r0 && col[k] > i; --k) ++nup;
    const int nfl = degU[i] - nup;                  // foreign lowers of row i
    for (int k = r0; k < r1; ++k) {
        const int j = col[k];
        if (k - r0 < nfl) {
            pid[k] = ustart[i] + (k - r0);
        } else if (j > i) {
            pid[k] = ustart[i] + nfl + (k - (r1 - nup));
        } else {
            const int jr1 = rowptr[j + 1];
            int lo = rowptr[j], hi = jr1;           // first entry of row j with col > j
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] > j) hi = mid; else lo = mid + 1; }
            const int fu = lo;
            const int nflj = degU[j] - (jr1 - fu);
            hi = jr1 - 1;                           // uppers of row j contain i
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < i) lo = mid + 1; else hi = mid; }
            pid[k] = ustart[j] + nflj + (lo - fu);
        }
    }
}

// e[p][k] and near[p] for every unordered pair; one thread per pair, 128 pairs per block.
//
// The 48 Gaussians sit on a uniform grid mu_k = mu_0 + k*dmu, so with t = D - mu_0
//     g_k = exp(-2 (t - k dmu)^2),   g_{k+1} = g_k * rho_k,   rho_{k+1} = rho_k * exp(-4 dmu^2),   rho_0 = exp(4 t dmu - 2 dmu^2)
// i.e. two exp() and 96 multiplications per pair instead of 48 exp() (float64 throughout: the recurrence's
// rounding error, ~1e-14 relative, is eight orders below the float32 rounding of e).  The three centres around D --
// which hold max_k e_k, the quantity the reference's is_near predicate tests (charge_gn.py:90-94) -- are evaluated
// with the reference's own formula, so the near flag and those entries are bit-exact.
// Rows are staged in shared memory and written as whole 128-byte lines.
// EKOUT == ED : the 48 float32 descriptor values (FP64 verification path).
// EKOUT == EDR: their EDR coefficients c = B^T e in the orthonormal basis B (float64 dot products of the float32-rounded
//               e, rounded once to float32): what the FP32 pair kernels consume (C^T e = (B^T C)^T c up to 5e-10).
#define EDGE_PAIRS 128
// gather_i != NULL: the distance is not read from pair_D but evaluated here from the pair's coordinates (fused list building of
// small-system chunks, epnn_bundle_prep.cu: one thread per pair, no divergence around the float64 arithmetic).
template <int EKOUT>
__global__ void __launch_bounds__(EDGE_PAIRS) edge_desc_kernel(int64_t P, const double* __restrict__ pair_D,
                                                               float* __restrict__ e, unsigned char* __restrict__ near,
                                                               unsigned long long* __restrict__ near_count,
                                                               const int* __restrict__ gather_i, const int* __restrict__ gather_j,
                                                               const float* __restrict__ xyz) {
    __shared__ float tile[EDGE_PAIRS][ED + 1];
    const int64_t p0 = (int64_t)blockIdx.x * EDGE_PAIRS;
    const int64_t p = p0 + threadIdx.x;
    bool is_near = false;
    if (p < P) {
        double D;
        if (gather_i) {
            const int i = gather_i[p], j = gather_j[p];
            D = dist64(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]);
        } else D = pair_D[p];
        const double C = cutoff_fn(D);
        const double dmu = c_mu[1] - c_mu[0];
        const double t = D - c_mu[0];
        double g = exp(-2.0 * t * t);
        double rho = exp(4.0 * t * dmu - 2.0 * dmu * dmu);
        const double q = exp(-4.0 * dmu * dmu);
        float* row = tile[threadIdx.x];
#pragma unroll 8
        for (int k = 0; k < ED; ++k) {
            row[k] = __double2float_rn(C * g);
            g *= rho;
            rho *= q;
        }
        int kc = (int)floor(t / dmu + 0.5);
        kc = min(ED - 2, max(1, kc));
        float emax = 0.f;
#pragma unroll
        for (int dk = -1; dk <= 1; ++dk) {
            const float ef = rbf_value(C, D, kc + dk);       // the reference's arithmetic, bit for bit
            row[kc + dk] = ef;
            emax = fmaxf(emax, ef);
        }
        is_near = emax > 1e-5f;
        near[p] = is_near ? 1 : 0;
        if (EKOUT != ED) {
            double cr[EDR];
#pragma unroll
            for (int r = 0; r < EDR; ++r) cr[r] = 0.0;
            for (int k = 0; k < ED; ++k) {
                const double ek = (double)row[k];
#pragma unroll
                for (int r = 0; r < EDR; ++r) cr[r] = fma(c_B[k * EDR + r], ek, cr[r]);
            }
#pragma unroll
            for (int r = 0; r < EDR; ++r) row[r] = __double2float_rn(cr[r]);
        }
    }
    if (near_count) {                                   // statistics (integer: order-independent): near pairs of this block
        const unsigned b = __ballot_sync(0xffffffffu, is_near);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(near_count, (unsigned long long)__popc(b));
    }
    __syncthreads();
    const int rows = (int)min((int64_t)EDGE_PAIRS, P - p0);
    float* dst = e + p0 * EKOUT;
    for (int f = threadIdx.x; f < rows * EKOUT; f += EDGE_PAIRS) dst[f] = tile[f / EKOUT][f % EKOUT];
}

#ifndef EPNN_CPU_EMU
cudaError_t launch_nbr_fill(const Workspace& w, const CellWork& cw, cudaStream_t st, int* nl) {
    if (w.n_atoms == 0) return cudaSuccess;
    nbr_kernel<true><<<div_up(w.n_atoms, NBR_BLOCK), NBR_BLOCK, 0, st>>>(w.n_atoms, w.atom_sys, w.sys_off, w.xyz, nullptr, nullptr,
                                                             w.rowptr, w.ustart, w.col, w.pair_i, w.pair_j, w.pair_D,
                                                             cw.grid, cw.cell_start, cw.cell_atoms, cw.Dtmp, w.row_lo, w.row_hi);
    ++*nl;
    nbr_rev_kernel<<<div_up(w.n_atoms, 128), 128, 0, st>>>(w.n_atoms, w.rowptr, w.ustart, w.degU, w.col, w.pid);
    ++*nl;
    return launch_edge_desc(w, st, nl, false);
}

cudaError_t launch_edge_desc(const Workspace& w, cudaStream_t st, int* nl, bool gather) {
    if (w.P > 0) {
        const int* gi = gather ? w.pair_i : nullptr;
        if (w.ek == ED) edge_desc_kernel<ED><<<div_up(w.P, EDGE_PAIRS), EDGE_PAIRS, 0, st>>>(w.P, w.pair_D, w.e, w.near, w.near_counter, gi, w.pair_j, w.xyz);
        else            edge_desc_kernel<EDR><<<div_up(w.P, EDGE_PAIRS), EDGE_PAIRS, 0, st>>>(w.P, w.pair_D, w.e, w.near, w.near_counter, gi, w.pair_j, w.xyz);
        ++*nl;
    }
    return cudaGetLastError();
}
#endif

// ------------------------------------------------------------------------------------------------
// Dense (n,n,48) descriptors of one system, literal layout of get_init_edges (facade / tests).
__global__ void edges_dense_kernel(int n, const float* __restrict__ xyz, float* __restrict__ e) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t tot = (int64_t)n * n * ED;
    if (idx >= tot) return;
    const int k = (int)(idx % ED);
    const int64_t ij = idx / ED;
    const int i = (int)(ij / n), j = (int)(ij % n);
    float out = 0.f;
    if (i != j) {
        const double D = dist64(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]);
        const double C = cutoff_fn(D);
        out = rbf_value(C, D, k);
    }
    e[idx] = out;
}

#ifndef EPNN_CPU_EMU
cudaError_t launch_edges_dense(int n, const float* xyz, float* e, cudaStream_t st) {
    const int64_t tot = (int64_t)n * n * ED;
    if (tot == 0) return cudaSuccess;
    edges_dense_kernel<<<div_up(tot, 256), 256, 0, st>>>(n, xyz, e);
    return cudaGetLastError();
}
#endif

// ------------------------------------------------------------------------------------------------
// Exclusive scan of int32 (out has n+1 entries; out[n] = total).  Three small kernels:
// per-block (1024 elements) sums -> single-block scan of the block sums -> per-block rescan + offset.
#define SCAN_BLOCK 1024

__global__ void scan_block_sums(const int* __restrict__ in, int n, int* __restrict__ bsum) {
    __shared__ int sh[32];
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    int v = i < n ? in[i] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = sh[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) bsum[blockIdx.x] = t;
    }
}

__global__ void scan_single(int* __restrict__ bsum, int nb, int* __restrict__ total) {
    // one block of 1024 threads scans nb values in strips
    __shared__ int sh[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? bsum[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = sh[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (threadIdx.x >= o) t += y; }
            sh[threadIdx.x] = t;
        }
        __syncthreads();
        const int warp_prefix = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
        const int incl = x + warp_prefix + carry;
        if (i < nb) bsum[i] = incl - v;           // exclusive
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void scan_apply(const int* __restrict__ in, int n, const int* __restrict__ boff, int* __restrict__ out) {
    __shared__ int sh[32];
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = sh[threadIdx.x];
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (threadIdx.x >= o) t += y; }
        sh[threadIdx.x] = t;
    }
    __syncthreads();
    const int warp_prefix = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
    if (i < n) out[i] = x - v + warp_prefix + boff[blockIdx.x];
}

// tmp needs div_up(n, 1024) ints.  out[n] receives the total.
#ifndef EPNN_CPU_EMU
cudaError_t launch_scan_i32(const int* in, int* out, int n, int* tmp, cudaStream_t st, int* nl) {
    if (n == 0) return cudaMemsetAsync(out, 0, sizeof(int), st);
    const int nb = div_up(n, SCAN_BLOCK);
    scan_block_sums<<<nb, SCAN_BLOCK, 0, st>>>(in, n, tmp);
    scan_single<<<1, SCAN_BLOCK, 0, st>>>(tmp, nb, out + n);
    scan_apply<<<nb, SCAN_BLOCK, 0, st>>>(in, n, tmp, out);
    *nl += 3;
    return cudaGetLastError();
}
#endif
