// CPU emulation driver for the DEFAULT bundle kernels and their list-building kernels, epnn_b200/csrc/epnn_bundle.cu
// (test infrastructure; see cuda_emu.h).  The far lists, the species-compressed far lists and the tile permutation are
// produced by the real prep kernels (far_count / bundle_mark / far_fill / far0 / tile_perm), run thread by thread.
// Build: g++ -O1 -std=c++17 -shared -fPIC -pthread -DEPNN_CPU_EMU -o build/libemu_bundle.so tools/emu/emu_bundle.cpp
#define EPNN_CPU_EMU 1
#include "../../epnn_b200/csrc/epnn_bundle.cu"

static void exclusive_scan(const int* in, int* out, int n) { int s = 0; for (int i = 0; i < n; ++i) { out[i] = s; s += in[i]; } out[n] = s; }

// Lists for one chunk.  Outputs must be sized by the caller: far_off/far0_off/rep/atom_b0/bundle_nat [n_atoms + 1],
// far_list [sum of (n_sys + 1) over atoms], far0_list/far0_w [17 * n_atoms], perm_j [P].  Returns the two list lengths.
extern "C" int emu_bundle_lists(int n_atoms, int n_bundles, const int* bundle_xy, int P,
                                const int* atom_sys, const int* sys_off, const int* npad, const int* species,
                                const int* rowptr, const int* col, const int* ustart, const int* pair_i, const int* pair_j,
                                int* atom_b0, int* bundle_nat, int* far_off, unsigned short* far_list,
                                int* far0_off, unsigned short* far0_list, unsigned char* far0_w, int* rep, unsigned char* perm_j,
                                int* n_far_out, int* n_far0_out) {
    const int2* bundle = reinterpret_cast<const int2*>(bundle_xy);
    std::vector<int> cnt(n_atoms + 1, 0);
    emu_launch_simple(div_up(n_atoms, 256), 256, [&] { far_count_kernel(n_atoms, atom_sys, sys_off, npad, rowptr, cnt.data()); });
    emu_launch_simple(div_up(n_bundles, 128), 128, [&] { bundle_mark_kernel(n_bundles, bundle, atom_b0, bundle_nat); });
    exclusive_scan(cnt.data(), far_off, n_atoms);
    emu_launch_simple(div_up(n_atoms, 128), 128, [&] { far_fill_kernel(n_atoms, atom_sys, sys_off, npad, rowptr, col, atom_b0, far_off, far_list); });
    if (P > 0)
        emu_launch_simple(div_up(P, 256), 256, [&] { tile_perm_kernel((int64_t)P, pair_i, pair_j, atom_sys, sys_off, atom_b0, ustart, bundle_nat, perm_j); });
    std::vector<int> cnt0(n_atoms + 1, 0);
    emu_launch_simple(div_up(n_atoms, 128), 128, [&] {
        far0_kernel<0>(n_atoms, atom_sys, sys_off, npad, species, rowptr, col, nullptr, rep, cnt0.data(), nullptr, nullptr, nullptr); });
    exclusive_scan(cnt0.data(), far0_off, n_atoms);
    emu_launch_simple(div_up(n_atoms, 128), 128, [&] {
        far0_kernel<1>(n_atoms, atom_sys, sys_off, npad, species, rowptr, col, atom_b0, nullptr, nullptr, far0_off, far0_list, far0_w); });
    *n_far_out = far_off[n_atoms];
    *n_far0_out = far0_off[n_atoms];
    return 0;
}

// weights: Cw[16*32] | W2[32*32] | b2[32] | x32[32] (b1: GNN, w3: EPN) -- as device arrays in the real launch
extern "C" int emu_bundle_kernel(int epn, const float* weights, int n_bundles, const int* bundle_xy, int* work_counter,
                                 const int* ustart, const int* pair_i, const int* pair_j, const unsigned char* near, const float* e,
                                 const unsigned char* perm_j, const int* far_off, const unsigned short* far_list,
                                 const int* far0_off, const unsigned short* far0_list, const unsigned char* far0_w, const int* rep, int dedup,
                                 const int* atom_sys, const int* sys_off, const int* npad,
                                 const float* u, const float* v, float* S, float* delta) {
    BundleArgs<float> a;
    a.n_bundles = n_bundles; a.bundle = reinterpret_cast<const int2*>(bundle_xy); a.work_counter = work_counter;
    a.ustart = ustart; a.pair_i = pair_i; a.pair_j = pair_j; a.near = near; a.e = e; a.perm_j = perm_j;
    a.far_off = far_off; a.far_list = far_list;
    a.far0_off = far0_off; a.far0_list = far0_list; a.far0_w = far0_w; a.rep = rep; a.dedup = dedup;
    a.atom_sys = atom_sys; a.sys_off = sys_off; a.npad = npad;
    a.u = u; a.v = v;
    a.Cw = weights; a.W2 = weights + EDR * HID; a.b2 = weights + EDR * HID + HID * HID; a.x32 = weights + EDR * HID + HID * HID + HID;
    a.S = S; a.delta = delta;
    *work_counter = 0;
    constexpr int NW = 8;
    if (epn) emu_launch_cta(NW, BundleSmem<float, true>::bytes(NW) / sizeof(float) + 8, [&] { bundle_kernel<float, NW, true>(a); });
    else     emu_launch_cta(NW, BundleSmem<float, false>::bytes(NW) / sizeof(float) + 8, [&] { bundle_kernel<float, NW, false>(a); });
    return 0;
}
