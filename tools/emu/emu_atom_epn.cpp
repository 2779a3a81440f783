// CPU emulation driver for the per-atom kernel (epnn_atom.cu) and the large-system electron-passing pair kernel
// (epnn_epn.cu) -- test infrastructure; see cuda_emu.h.
// Build: g++ -O1 -std=c++17 -shared -fPIC -pthread -DEPNN_CPU_EMU -o build/libemu_atom_epn.so tools/emu/emu_atom_epn.cpp
#define EPNN_CPU_EMU 1
#include "../../epnn_b200/csrc/epnn_atom.cu"
#include "../../epnn_b200/csrc/epnn_epn.cu"
#include "../../epnn_b200/csrc/epnn_atom_const.cu"
#include "../../epnn_b200/csrc/epnn_atom_mma.cu"

// wu: HG[64*32] | cb[32] | g[32] | U2[32*32] | c2[32] | U3[32*48] | c3[48]        (update side, already folded)
// wp: Pf[32*64] | Aq64[64] | Ax[16*64]                                            (projection side of the NEXT pair kernel)
extern "C" int emu_atom_kernel(int mode, int h_is_zero, int n_atoms, int nsplit, const float* wu, const float* wp,
                               const int* atom_sys, const int* sys_off, const int* npad, const int* species,
                               const float* Spart, float* h, float* l2, const int* rowptr, const int* col, const int* pid,
                               const float* delta, double* q, float* u, float* v, float* q_out, double* q_out64) {
    AtomArgs<float, float> aa;
    memset(&aa, 0, sizeof(aa));
    aa.n_atoms = n_atoms; aa.mode = mode; aa.nsplit = nsplit; aa.h_is_zero = h_is_zero;
    aa.atom_sys = atom_sys; aa.sys_off = sys_off; aa.npad = npad; aa.species = species;
    aa.Spart = Spart; aa.h = h; aa.l2 = l2;
    aa.HG = wu; aa.cb = wu + 64 * HID; aa.g = aa.cb + HID;
    aa.upd.U2 = aa.g + HID; aa.upd.c2 = aa.upd.U2 + HID * HID; aa.upd.U3 = aa.upd.c2 + HID; aa.upd.c3 = aa.upd.U3 + HID * HD;
    aa.rowptr = rowptr; aa.col = col; aa.pid = pid; aa.delta = delta; aa.q = q;
    aa.Pf = wp; aa.Aq64 = wp + HID * 64; aa.Ax = aa.Aq64 + 64;
    aa.u = u; aa.v = v; aa.q_out = q_out; aa.q_out64 = q_out64;
    constexpr int NW = 4;
    const size_t smem = sizeof(float) * (ATOM_W_UPD + ATOM_W_PROJ + (size_t)NW * ATOM_TILE + NW * 64) + sizeof(int) * NW * 64;
    emu_launch_grid(2, NW, smem / sizeof(float) + 8, [&] { atom_kernel<float, NW, float, false>(aa); });
    return 0;
}

// the warp-level tensor variant (epnn_atom_mma.cu: FP32 default), same inputs
extern "C" int emu_atom_mma_kernel(int mode, int h_is_zero, int n_atoms, int nsplit, const float* wu, const float* wp,
                                   const int* atom_sys, const int* sys_off, const int* npad, const int* species,
                                   const float* Spart, float* h, float* l2, const int* rowptr, const int* col, const int* pid,
                                   const float* delta, double* q, float* u, float* v, float* q_out, double* q_out64) {
    AtomArgs<float, float> aa;
    memset(&aa, 0, sizeof(aa));
    aa.n_atoms = n_atoms; aa.mode = mode; aa.nsplit = nsplit; aa.h_is_zero = h_is_zero;
    aa.atom_sys = atom_sys; aa.sys_off = sys_off; aa.npad = npad; aa.species = species;
    aa.Spart = Spart; aa.h = h; aa.l2 = l2;
    aa.HG = wu; aa.cb = wu + 64 * HID; aa.g = aa.cb + HID;
    aa.upd.U2 = aa.g + HID; aa.upd.c2 = aa.upd.U2 + HID * HID; aa.upd.U3 = aa.upd.c2 + HID; aa.upd.c3 = aa.upd.U3 + HID * HD;
    aa.rowptr = rowptr; aa.col = col; aa.pid = pid; aa.delta = delta; aa.q = q;
    aa.Pf = wp; aa.Aq64 = wp + HID * 64; aa.Ax = aa.Aq64 + 64;
    aa.u = u; aa.v = v; aa.q_out = q_out; aa.q_out64 = q_out64;
    if (mode & ATOM_UPDATE) emu_launch_grid(2, AmL<true>::NW, AmL<true>::WORDS + 8, [&] { atom_mma_kernel<false, true>(aa); });
    else emu_launch_grid(2, AmL<false>::NW, AmL<false>::WORDS + 8, [&] { atom_mma_kernel<false, false>(aa); });
    return 0;
}

// the experimental atom-per-thread variant (epnn_atom_const.cu), same inputs
extern "C" int emu_atom_const_kernel(int mode, int h_is_zero, int n_atoms, int nsplit, const float* wu, const float* wp,
                                     const int* atom_sys, const int* sys_off, const int* npad, const int* species,
                                     const float* Spart, float* h, float* l2, const int* rowptr, const int* col, const int* pid,
                                     const float* delta, double* q, float* u, float* v, float* q_out, double* q_out64) {
    AtomW W;
    const float* p = wu;
    memcpy(W.HG, p, sizeof(W.HG)); p += 64 * HID; memcpy(W.cb, p, sizeof(W.cb)); p += HID; memcpy(W.g, p, sizeof(W.g)); p += HID;
    memcpy(W.U2, p, sizeof(W.U2)); p += HID * HID; memcpy(W.c2, p, sizeof(W.c2)); p += HID;
    memcpy(W.U3, p, sizeof(W.U3)); p += HID * HD; memcpy(W.c3, p, sizeof(W.c3));
    memcpy(W.Pf, wp, sizeof(W.Pf)); memcpy(W.Aq, wp + HID * 64, sizeof(W.Aq));
    AtomConstArgs aa;
    memset(&aa, 0, sizeof(aa));
    aa.n_atoms = n_atoms; aa.mode = mode; aa.nsplit = nsplit; aa.h_is_zero = h_is_zero;
    aa.atom_sys = atom_sys; aa.sys_off = sys_off; aa.npad = npad; aa.species = species;
    aa.Spart = Spart; aa.h = h; aa.l2 = l2; aa.rowptr = rowptr; aa.col = col; aa.pid = pid; aa.delta = delta; aa.q = q;
    aa.Ax = wp + HID * 64 + 64; aa.u = u; aa.v = v; aa.q_out = q_out; aa.q_out64 = q_out64;
    emu_launch_grid(2, ACONST_NW, (size_t)ACONST_NW * 32 * ATS, [&] { atom_const_kernel(W, aa); });
    return 0;
}

// weights: Cw[16*32] | W2[32*32] | b2[32] | w3[32]
extern "C" int emu_epn_pair_kernel(const float* weights, int P, const int* pair_i, const int* pair_j, const unsigned char* near,
                                   const float* e, const int* atom_sys, const int* sys_off, const float* u, const float* v, float* delta) {
    EpnArgs<float> ea;
    ea.P = P; ea.tile_begin = 0; ea.tile_end = (P + 31) / 32;
    ea.pair_i = pair_i; ea.pair_j = pair_j; ea.near = near; ea.e = e; ea.atom_sys = atom_sys; ea.sys_off = sys_off;
    ea.u = u; ea.v = v;
    ea.Cw = weights; ea.W2 = weights + EDR * HID; ea.b2 = weights + EDR * HID + HID * HID; ea.w3 = weights + EDR * HID + HID * HID + HID;
    ea.delta = delta;
    constexpr int NW = 8;
    const size_t smem = sizeof(float) * (EDR * HID + HID * HID + 2 * HID + (size_t)NW * (32 * EDR + 32 * HID)) + sizeof(int) * NW * 64;
    emu_launch_grid(2, NW, smem / sizeof(float) + 8, [&] { epn_pair_kernel<float, NW>(ea); });
    return 0;
}
