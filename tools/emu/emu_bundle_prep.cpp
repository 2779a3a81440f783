// CPU emulation driver for the fused list building of small-system chunks, epnn_b200/csrc/epnn_bundle_prep.cu (test
// infrastructure; see cuda_emu.h): count pass -> four scans over bundles (host) -> fill pass, as launch_bundle_prep_* do.
// Build: g++ -O1 -std=c++17 -ffp-contract=off -shared -fPIC -pthread -DEPNN_CPU_EMU -o build/libemu_bundle_prep.so tools/emu/emu_bundle_prep.cpp
#define EPNN_CPU_EMU 1
#include "../../epnn_b200/csrc/epnn_bundle_prep.cu"
#include "../../epnn_b200/csrc/epnn_neighbor.cu"        // edge_desc_kernel: distances + descriptors of the pair list (gather mode)

extern "C" int emu_bundle_prep(int n_bundles, int n_atoms, const int* bundle, const int* atom_sys, const int* sys_off, const int* npad,
                               const int* species, const float* xyz, int* totals /* nnz, P, far, far0 */,
                               int* deg, int* degU, int* rowptr, int* ustart, int* far_off, int* far0_off, int* rep, int* atom_b0, int* bundle_nat,
                               int* col, int* pid, unsigned char* rowl, int* pair_i, int* pair_j, const double* mu, const double* B, float* coef, unsigned char* near,
                               unsigned short* far_list, unsigned short* far0_list, unsigned char* far0_w, int cap) {
    std::vector<unsigned long long> mask((size_t)n_atoms);
    std::vector<int> btot(4 * (size_t)n_bundles), boff(4 * ((size_t)n_bundles + 1));
    BundlePrepArgs a;
    memset(&a, 0, sizeof(a));
    a.n_bundles = n_bundles; a.n_atoms = n_atoms; a.bundle = reinterpret_cast<const int2*>(bundle);
    a.atom_sys = atom_sys; a.sys_off = sys_off; a.npad = npad; a.species = species; a.xyz = xyz;
    a.mask = mask.data(); a.btot = btot.data(); a.boff = boff.data();
    emu_launch_grid(2, BP_NW, 0, [&] { bundle_count_kernel(a); });
    for (int k = 0; k < 4; ++k) {
        int s = 0;
        for (int b = 0; b < n_bundles; ++b) { boff[(size_t)k * (n_bundles + 1) + b] = s; s += btot[(size_t)k * n_bundles + b]; }
        boff[(size_t)k * (n_bundles + 1) + n_bundles] = s;
        totals[k] = s;
        if (s > cap) return -1;
    }
    a.deg = deg; a.degU = degU; a.rowptr = rowptr; a.ustart = ustart; a.far_off = far_off; a.far0_off = far0_off; a.rep = rep;
    a.atom_b0 = atom_b0; a.bundle_nat = bundle_nat; a.col = col; a.pid = pid; a.rowl = rowl;
    a.pair_i = pair_i; a.pair_j = pair_j; a.far_list = far_list; a.far0_list = far0_list; a.far0_w = far0_w;
    emu_launch_grid(2, BP_NW, 0, [&] { bundle_fill_kernel(a); });
    emu_set_rbf(mu, B);
    const int64_t P = totals[1];
    if (P > 0) emu_launch_grid(div_up(P, EDGE_PAIRS), EDGE_PAIRS / 32, 0, [&] { edge_desc_kernel<EDR>(P, nullptr, coef, near, nullptr, pair_i, pair_j, xyz); });
    return 0;
}
