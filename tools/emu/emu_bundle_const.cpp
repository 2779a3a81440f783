// CPU emulation driver for epnn_b200/csrc/epnn_bundle_const.cu (test infrastructure; see cuda_emu.h).
// Build: g++ -O1 -std=c++17 -shared -fPIC -pthread -DEPNN_CPU_EMU -o build/libemu_bundle_const.so tools/emu/emu_bundle_const.cpp
#define EPNN_CPU_EMU 1
#include "../../epnn_b200/csrc/epnn_bundle_const.cu"

// weights: Cw[16*32] | W2[32*32] | b2[32] | x32[32] (b1 for the GNN variant, w3 for the EPN variant)
extern "C" int emu_bundle_const(int epn, int n_warps, const float* weights,
                                int n_bundles, const int* bundle_xy, int* work_counter,
                                const int* ustart, const int* pair_i, const int* pair_j, const unsigned char* near, const float* e,
                                const int* far_off, const unsigned short* far_list,
                                const int* far0_off, const unsigned short* far0_list, const unsigned char* far0_w, const int* rep, int dedup,
                                const int* atom_sys, const int* sys_off, const int* npad,
                                const float* u, const float* v, float* S, float* delta) {
    PairW W;
    memcpy(W.W2, weights + EDR * HID, sizeof(W.W2));
    ConstArgs a;
    a.n_bundles = n_bundles; a.bundle = reinterpret_cast<const int2*>(bundle_xy); a.work_counter = work_counter;
    a.ustart = ustart; a.pair_i = pair_i; a.pair_j = pair_j; a.near = near; a.e = e;
    a.far_off = far_off; a.far_list = far_list;
    a.far0_off = far0_off; a.far0_list = far0_list; a.far0_w = far0_w; a.rep = rep; a.dedup = dedup;
    a.atom_sys = atom_sys; a.sys_off = sys_off; a.npad = npad;
    a.u = u; a.v = v; a.S = S; a.delta = delta;
    a.Cw = weights; a.b2 = weights + EDR * HID + HID * HID; a.x32 = weights + EDR * HID + HID * HID + HID;
    *work_counter = 0;
    if (n_warps < 1 || n_warps > CONST_NW) return -1;
    if (epn) emu_launch_cta(n_warps, (size_t)CONST_NW * ConstSmem<true>::PW + ConstSmem<true>::SHARED, [&] { bundle_const_kernel<true>(W, &a); });
    else     emu_launch_cta(n_warps, (size_t)CONST_NW * ConstSmem<false>::PW + ConstSmem<false>::SHARED, [&] { bundle_const_kernel<false>(W, &a); });
    return 0;
}
