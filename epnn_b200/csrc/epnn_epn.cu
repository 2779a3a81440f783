// Electron-passing pair kernel of the EPN layer.
//
// Replaces, per pass t, reference charge_gn.py:101-116 (EPN_layer.call): two (N*N, K) pair-input
// tensors, pass_fns[t] on all N^2 pairs twice, antisymmetrise and mask by is_near.  Here each UNORDERED
// pair of the e != 0 list is evaluated once, both directions in the same tile (they share C^T e_ij):
//     f_ij  = w3 . relu(W2^T relu(u_i + v_j + C^T e_ij) + b2)
//     f_ji  = w3 . relu(W2^T relu(u_j + v_i + C^T e_ij) + b2)          (final bias cancels)
//     delta_p = 0.5 (f_ij - f_ji) * is_near_p
// The +delta / -delta scatter into q is the per-atom kernel's fixed-order row reduction, so charge is
// conserved by construction and the result does not depend on scheduling.
//
// Work unit = one warp on a tile of 32 consecutive pairs; thread (pg, og) owns pairs pg*8..pg*8+7 and
// hidden columns og*4..og*4+3.
#include "epnn_internal.cuh"

template <typename R> struct EpnArgs {
    int64_t P; int64_t tile_begin, tile_end;
    const int* pair_i; const int* pair_j; const unsigned char* near; const float* e;
    const int* atom_sys; const int* sys_off;
    const R* u; const R* v;
    const R* Cw; const R* W2; const R* b2; const R* w3;
    R* delta;
};

template <typename R, int NW>
__global__ void __launch_bounds__(NW * 32, sizeof(R) == 4 ? 2 : 1) epn_pair_kernel(const EpnArgs<R> a) {
#ifdef EPNN_CPU_EMU
    unsigned char* smem_raw = reinterpret_cast<unsigned char*>(emu_smem);
#else
    extern __shared__ __align__(32) unsigned char smem_raw[];
#endif
    constexpr int EK = EKof<R>::v;
    R* sC = reinterpret_cast<R*>(smem_raw);          // [48][32]
    R* sW2 = sC + EK * HID;                          // [32][32]
    R* sb2 = sW2 + HID * HID;                        // [32]
    R* sw3 = sb2 + HID;                              // [32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    R* at1 = sw3 + HID + warp * (32 * EK + 32 * HID);
    R* at2 = at1 + 32 * EK;
    int* slot_i = reinterpret_cast<int*>(sw3 + HID + NW * (32 * EK + 32 * HID)) + warp * 64;
    int* slot_j = slot_i + 32;

    for (int t = threadIdx.x; t < EK * HID; t += NW * 32) sC[t] = a.Cw[t];
    for (int t = threadIdx.x; t < HID * HID; t += NW * 32) sW2[t] = a.W2[t];
    if (threadIdx.x < HID) { sb2[threadIdx.x] = a.b2[threadIdx.x]; sw3[threadIdx.x] = a.w3[threadIdx.x]; }
    __syncthreads();

    const int pg = lane >> 3, og = lane & 7;
    const Vec4<R> b2v = ldv(sb2 + og * 4);
    const Vec4<R> w3v = ldv(sw3 + og * 4);
    for (int64_t tile = a.tile_begin + (int64_t)blockIdx.x * NW + warp; tile < a.tile_end; tile += (int64_t)gridDim.x * NW) {
        const int64_t p = tile * 32 + lane;
        bool valid = p < a.P;
        int pi = -1, pj = -1;
        if (valid) {                                  // pairs of small systems belong to the bundle kernel
            pi = a.pair_i[p]; pj = a.pair_j[p];
            const int sys = a.atom_sys[pi];
            valid = a.sys_off[sys + 1] - a.sys_off[sys] > SMALL_MAX;
        }
        if (!__any_sync(0xffffffffu, valid)) continue;
        slot_i[lane] = valid ? pi : -1;
        slot_j[lane] = valid ? pj : -1;
        const R nearf = valid ? (R)a.near[p] : R(0);
        const int64_t rows_left = a.P - tile * 32;
        const float4* esrc = reinterpret_cast<const float4*>(a.e + tile * 32 * EK);
#pragma unroll 4
        for (int f = lane; f < 32 * (EK / 4); f += 32) {           // 32 descriptor rows, contiguous in HBM
            const int sl = f / (EK / 4), ch = f - sl * (EK / 4);
            Vec4<R> ev = vzero<R>();
            if (sl < rows_left) ev = cvt4<R>(__ldg(esrc + f));
            stv(at1 + tile_off(sl, ch, EK), ev);
        }
        __syncwarp();
        R ce[8][4];
        zero_acc(ce);
        tile_gemm<R, EK, HID>(at1, sC, og * 4, ce, pg);

        R part[8];
        R acc[8][4];
        // ---- direction i <- j
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int ii = slot_i[pg * 8 + s], jj = slot_j[pg * 8 + s];
            Vec4<R> z = vzero<R>();
            if (ii >= 0) {
                const Vec4<R> ui = ldv(a.u + (int64_t)ii * HID + og * 4);
                const Vec4<R> vj = ldv(a.v + (int64_t)jj * HID + og * 4);
                z.x = relu(ce[s][0] + ui.x + vj.x); z.y = relu(ce[s][1] + ui.y + vj.y);
                z.z = relu(ce[s][2] + ui.z + vj.z); z.w = relu(ce[s][3] + ui.w + vj.w);
            }
            stv(at2 + tile_off(pg * 8 + s, og, HID), z);
        }
        __syncwarp();
        zero_acc(acc);
        tile_gemm<R, HID, HID>(at2, sW2, og * 4, acc, pg);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            R f = relu(acc[s][0] + b2v.x) * w3v.x;
            f = fma(relu(acc[s][1] + b2v.y), w3v.y, f);
            f = fma(relu(acc[s][2] + b2v.z), w3v.z, f);
            f = fma(relu(acc[s][3] + b2v.w), w3v.w, f);
            part[s] = f;
        }
        __syncwarp();
        // ---- direction j <- i
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int ii = slot_i[pg * 8 + s], jj = slot_j[pg * 8 + s];
            Vec4<R> z = vzero<R>();
            if (ii >= 0) {
                const Vec4<R> uj = ldv(a.u + (int64_t)jj * HID + og * 4);
                const Vec4<R> vi = ldv(a.v + (int64_t)ii * HID + og * 4);
                z.x = relu(ce[s][0] + uj.x + vi.x); z.y = relu(ce[s][1] + uj.y + vi.y);
                z.z = relu(ce[s][2] + uj.z + vi.z); z.w = relu(ce[s][3] + uj.w + vi.w);
            }
            stv(at2 + tile_off(pg * 8 + s, og, HID), z);
        }
        __syncwarp();
        zero_acc(acc);
        tile_gemm<R, HID, HID>(at2, sW2, og * 4, acc, pg);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            R f = relu(acc[s][0] + b2v.x) * w3v.x;
            f = fma(relu(acc[s][1] + b2v.y), w3v.y, f);
            f = fma(relu(acc[s][2] + b2v.z), w3v.z, f);
            f = fma(relu(acc[s][3] + b2v.w), w3v.w, f);
            part[s] -= f;
        }
        // ---- reduce-scatter part[0..7] over the 8 og lanes: lane og ends with the total of slot s = og
        R r4[4], r2[2], r1;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const R send = (og & 4) ? part[t] : part[t + 4];
            const R keep = (og & 4) ? part[t + 4] : part[t];
            r4[t] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const R send = (og & 2) ? r4[t] : r4[t + 2];
            const R keep = (og & 2) ? r4[t + 2] : r4[t];
            r2[t] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        {
            const R send = (og & 1) ? r2[0] : r2[1];
            const R keep = (og & 1) ? r2[1] : r2[0];
            r1 = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        if (valid) a.delta[p] = R(0.5) * r1 * nearf;       // charge_gn.py:116
        __syncwarp();
    }
}

#ifndef EPNN_CPU_EMU
template <typename R>
cudaError_t launch_epn_pair(const Workspace& w, const StepW<R>& sw, cudaStream_t st, int* nl) {
    if (w.P == 0 || w.n_rg_large == 0) return cudaSuccess;
    constexpr int NW = 8;
    EpnArgs<R> ea;
    ea.P = w.P; ea.pair_i = w.pair_i; ea.pair_j = w.pair_j; ea.near = w.near; ea.e = w.e; ea.atom_sys = w.atom_sys; ea.sys_off = w.sys_off;
    ea.u = (const R*)w.u; ea.v = (const R*)w.v; ea.Cw = sw.Cw; ea.W2 = sw.W2; ea.b2 = sw.b2; ea.w3 = sw.W3;
    ea.delta = (R*)w.delta;
    constexpr int EK = EKof<R>::v;
    const size_t smem = sizeof(R) * (EK * HID + HID * HID + 2 * HID + (size_t)NW * (32 * EK + 32 * HID)) + sizeof(int) * NW * 64;
    cudaError_t e = cudaFuncSetAttribute(epn_pair_kernel<R, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_sm = sizeof(R) == 4 ? 2 : 1;
    const int64_t n_tiles = (w.P + 31) / 32;
    ea.tile_begin = 0; ea.tile_end = n_tiles;      // sharded call: the local pair list already holds only the pairs incident to owned rows
    int64_t grid = (ea.tile_end - ea.tile_begin + NW - 1) / NW;
    if (grid < 1) return cudaSuccess;
    if (grid > (int64_t)w.sm_count * per_sm) grid = (int64_t)w.sm_count * per_sm;
    epn_pair_kernel<R, NW><<<(int)grid, NW * 32, smem, st>>>(ea);
    ++*nl;
    return cudaGetLastError();
}

template cudaError_t launch_epn_pair<float>(const Workspace&, const StepW<float>&, cudaStream_t, int*);
template cudaError_t launch_epn_pair<double>(const Workspace&, const StepW<double>&, cudaStream_t, int*);
#endif
