// CPU emulation driver for the large-system message kernel and its species tables, epnn_b200/csrc/epnn_gnn.cu
// (test infrastructure; see cuda_emu.h).
// Build: g++ -O1 -std=c++17 -shared -fPIC -pthread -DEPNN_CPU_EMU -o build/libemu_gnn.so tools/emu/emu_gnn.cpp
#define EPNN_CPU_EMU 1
#include "../../epnn_b200/csrc/epnn_gnn.cu"

// One message-passing step of the large systems of a chunk.  stamp == 0: de-duplication off.
// weights: Cw[16*32] | W2[32*32] | b2[32] | b1[32].  S: [nsplit][n_atoms][32] (zero-filled here).
// sp_tab [n_sp_tab * 32], sp_stamp [n_sp_tab * 2], dedup_rows [1] are scratch / outputs.
extern "C" int emu_gnn_large_step(const float* weights, int n_atoms, int n_sys, int n_rg, const int* rg_atom, int nsplit, int skip_far,
                                  const int* atom_sys, const int* sys_off, const int* npad, const int* species, const int* rgl_off,
                                  const int* rowptr, const int* col, const int* pid, const int* deg, const float* e,
                                  const float* u, const float* v, int stamp, int n_species, int n_sp_tab, int* sp_tab, int* sp_stamp,
                                  unsigned long long* dedup_rows, float* S, int shard_rank, int shard_world) {
    if (stamp) {                      // species tables + this step's equality check, exactly as launch_sp_tab_build / launch_sp_check
        emu_launch_simple(div_up((int64_t)n_sp_tab * 32, 256), 256, [&] { sp_tab_init_kernel(n_sp_tab, sp_tab, sp_stamp); });
        emu_launch_grid(div_up(n_atoms, 256), 8, 0, [&] { sp_tab_fill_kernel(n_atoms, atom_sys, sys_off, species, rgl_off, deg, sp_tab, sp_stamp); });
        emu_launch_simple(div_up((int64_t)n_atoms * 8, 256), 256, [&] { sp_check_kernel<float>(n_atoms, atom_sys, sys_off, species, rgl_off, sp_tab, v, sp_stamp, stamp); });
        *dedup_rows = 0;
        emu_launch_simple(div_up(n_sp_tab, 256), 256, [&] { sp_tally_kernel(n_sp_tab, sp_tab, sp_stamp, stamp, dedup_rows); });
    }
    GnnArgs<float> ga;
    // this rank's contiguous slice of the work units, as launch_gnn_pair computes it (world 1: everything)
    const int64_t total = (int64_t)n_rg * nsplit;
    ga.rg_atom = rg_atom; ga.unit_begin = (int)(total * shard_rank / shard_world); ga.n_units = (int)(total * (shard_rank + 1) / shard_world);
    ga.nsplit = nsplit; ga.n_atoms = n_atoms;
    ga.skip_far = skip_far; ga.plane = 0;
    ga.atom_sys = atom_sys; ga.sys_off = sys_off; ga.npad = npad; ga.rowptr = rowptr; ga.col = col; ga.pid = pid; ga.e = e;
    ga.u = u; ga.v = v;
    ga.Cw = weights; ga.W2 = weights + EDR * HID; ga.b2 = weights + EDR * HID + HID * HID; ga.b1 = weights + EDR * HID + HID * HID + HID;
    ga.S = S;
    ga.species = species; ga.rgl_off = rgl_off; ga.sp_tab = sp_tab; ga.sp_stamp = sp_stamp; ga.stamp = stamp; ga.n_species = n_species;
    memset(S, 0, sizeof(float) * (size_t)nsplit * n_atoms * HID);
    constexpr int NW = 8;
    emu_launch_grid(2, NW, gnn_smem_bytes<float>(NW) / sizeof(float) + 8, [&] { gnn_pair_kernel<float, true, NW>(ga); });
    return 0;
}
