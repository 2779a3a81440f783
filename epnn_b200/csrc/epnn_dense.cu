// Keras-shaped compatibility path: the reference model called on the dense padded tensors that
// charge_gn.gen_padded_init_state produces (reference charge_gn.py:369-391), evaluated LITERALLY:
//   un-tiling   h,x,q = reduce_sum(inp, axis=1) / reduce_sum(mask, axis=1)  (divide_no_nan)        :382-384
//   GNN_layer   all N*N pairs (padded ones included), unmasked reduce_sum over j, node_mask           :56-75
//   EPN_layer   f_ij, f_ji on all N*N pairs, 0.5 (f_ij - f_ji) * mask_ij * is_near(e_ij), sum over j :87-119
// Arbitrary x / h / q / mask / e are accepted (nothing is assumed about one-hot features, symmetry of e or the
// shape of the mask).  This is the drop-in for `model([h, e, x, q, mask])`; the packed fast path is
// epnn_infer_batch.  One CTA per (system, atom row); plain FP32 (or FP64) SIMT, FP64 reductions.
#include "epnn_internal.cuh"

#define DENSE_THREADS 128

// a[b][k][F] = [x | h | q] of atom k, node_mask[b][k] = clip(sum_j mask[b][j][k], 0, 1)
template <typename R>
__global__ void dense_untile_kernel(int N, int n_x, const float* __restrict__ h, const float* __restrict__ x,
                                    const float* __restrict__ q, const float* __restrict__ mask, R* __restrict__ a,
                                    R* __restrict__ node_mask) {
    const int b = blockIdx.x / N, k = blockIdx.x % N;
    const int F = n_x + HD + 1;
    __shared__ double den_s;
    if (threadIdx.x == 0) {
        double den = 0.0;
        for (int j = 0; j < N; ++j) den += (double)mask[((size_t)b * N + j) * N + k];
        den_s = den;
        node_mask[(size_t)b * N + k] = (R)fmin(fmax(den, 0.0), 1.0);
    }
    __syncthreads();
    const double den = den_s;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) {
            const size_t jk = ((size_t)b * N + j) * N + k;
            s += f < n_x ? (double)x[jk * n_x + f] : f < n_x + HD ? (double)h[jk * HD + (f - n_x)] : (double)q[jk];
        }
        a[((size_t)b * N + k) * F + f] = den != 0.0 ? (R)(s / den) : R(0);
    }
}

// u[b][k][0..31] = a_k . W1[a_i rows],  v[b][k][0..31] = a_k . W1[a_j rows] + b1        (64 threads)
template <typename R>
__global__ void dense_project_kernel(int n_x, const R* __restrict__ a, DenseW<R> w, R* __restrict__ uv) {
    const int F = n_x + HD + 1;
    const R* ak = a + (size_t)blockIdx.x * F;
    const int c = threadIdx.x;
    R s = c >= HID ? w.b1[c - HID] : R(0);
    for (int f = 0; f < n_x; ++f) s = fma(ak[f], w.Wx64[f * 64 + c], s);
    for (int f = 0; f < HD; ++f) s = fma(ak[n_x + f], w.Ah64[f * 64 + c], s);
    s = fma(ak[F - 1], w.Aq64[c], s);
    uv[(size_t)blockIdx.x * 64 + c] = s;
}

template <typename R>
__device__ __forceinline__ void hidden2(const R* __restrict__ sC, const R* __restrict__ sW2, const R* __restrict__ sb2,
                                        const R (&ce)[HID], const R* __restrict__ ui, const R* __restrict__ vj, R (&z2)[HID]) {
    R z1[HID];
#pragma unroll
    for (int c = 0; c < HID; ++c) z1[c] = relu(ce[c] + ui[c] + vj[c]);
#pragma unroll
    for (int c = 0; c < HID; ++c) z2[c] = sb2[c];
#pragma unroll 4
    for (int k = 0; k < HID; ++k) {
        const R zk = z1[k];
#pragma unroll
        for (int c = 0; c < HID; ++c) z2[c] = fma(zk, sW2[k * HID + c], z2[c]);
    }
#pragma unroll
    for (int c = 0; c < HID; ++c) z2[c] = relu(z2[c]);
}

// MODE 0: message step  -> msg[b][i][32] = W3^T sum_j relu(...) + N b3      (charge_gn.py:63-70)
// MODE 1: pass step     -> a[b][i][F-1] += sum_j 0.5 (f_ij - f_ji) mask_ij near_ij   (charge_gn.py:101-118)
template <typename R, int MODE>
__global__ void __launch_bounds__(DENSE_THREADS) dense_pair_kernel(int N, int n_x, const float* __restrict__ e,
                                                                  const float* __restrict__ mask, const R* __restrict__ uv,
                                                                  DenseW<R> w, R* __restrict__ msg, R* __restrict__ a) {
    __shared__ R sC[ED * HID];
    __shared__ R sW2[HID * HID];
    __shared__ R sb2[HID];
    __shared__ R sw3[HID * HID];
    __shared__ R sui[HID], svi[HID];
    __shared__ double red[DENSE_THREADS / 32][HID];
    const int b = blockIdx.x / N, i = blockIdx.x % N;
    const int F = n_x + HD + 1;
    for (int t = threadIdx.x; t < ED * HID; t += DENSE_THREADS) sC[t] = w.Cw[t];
    for (int t = threadIdx.x; t < HID * HID; t += DENSE_THREADS) sW2[t] = w.W2[t];
    for (int t = threadIdx.x; t < (MODE == 0 ? HID * HID : HID); t += DENSE_THREADS) sw3[t] = w.W3[t];
    if (threadIdx.x < HID) {
        sb2[threadIdx.x] = w.b2[threadIdx.x];
        sui[threadIdx.x] = uv[((size_t)b * N + i) * 64 + threadIdx.x];
        svi[threadIdx.x] = uv[((size_t)b * N + i) * 64 + HID + threadIdx.x];
    }
    __syncthreads();
    double accum[MODE == 0 ? HID : 1];
#pragma unroll
    for (int c = 0; c < (MODE == 0 ? HID : 1); ++c) accum[c] = 0.0;
    for (int j = threadIdx.x; j < N; j += DENSE_THREADS) {
        const float* eij = e + (((size_t)b * N + i) * N + j) * ED;
        R ce[HID];
#pragma unroll
        for (int c = 0; c < HID; ++c) ce[c] = R(0);
        float emax = 1e-5f;                                    // clip(e, 1e-5, 1e5).max()
        for (int k = 0; k < ED; ++k) {
            const float ek = eij[k];
            emax = fmaxf(emax, fminf(fmaxf(ek, 1e-5f), 1e5f));
            const R er = (R)ek;
#pragma unroll
            for (int c = 0; c < HID; ++c) ce[c] = fma(er, sC[k * HID + c], ce[c]);
        }
        const R* uvj = uv + ((size_t)b * N + j) * 64;
        R uj[HID], vj[HID];
#pragma unroll
        for (int c = 0; c < HID; ++c) { uj[c] = uvj[c]; vj[c] = uvj[HID + c]; }
        R z2[HID];
        hidden2<R>(sC, sW2, sb2, ce, sui, vj, z2);             // [a_i | a_j | e_ij]
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < HID; ++c) accum[c] += (double)z2[c];
        } else {
            R fij = w.b3[0];
#pragma unroll
            for (int c = 0; c < HID; ++c) fij = fma(z2[c], sw3[c], fij);
            hidden2<R>(sC, sW2, sb2, ce, uj, svi, z2);         // [a_j | a_i | e_ij]
            R fji = w.b3[0];
#pragma unroll
            for (int c = 0; c < HID; ++c) fji = fma(z2[c], sw3[c], fji);
            const R near = emax != 1e-5f ? R(1) : R(0);        // tf.not_equal(largest_e, tol), charge_gn.py:93
            const R m = (R)mask[((size_t)b * N + i) * N + j];
            accum[0] += (double)(R(0.5) * (fij - fji) * m * near);
        }
    }
    // block reduction in a fixed order: lanes (xor tree), then warps 0..3
    constexpr int NC = MODE == 0 ? HID : 1;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double v = accum[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (MODE == 0) {
        if (threadIdx.x < HID) {
            double S[HID];
            for (int k = 0; k < HID; ++k) { double t = 0.0; for (int wp = 0; wp < DENSE_THREADS / 32; ++wp) t += red[wp][k]; S[k] = t; }
            const int c = threadIdx.x;
            double m = (double)N * (double)w.b3[c];
            for (int k = 0; k < HID; ++k) m += S[k] * (double)sw3[k * HID + c];
            msg[((size_t)b * N + i) * HID + c] = (R)m;
        }
    } else if (threadIdx.x == 0) {
        double t = 0.0;
        for (int wp = 0; wp < DENSE_THREADS / 32; ++wp) t += red[wp][0];
        a[((size_t)b * N + i) * F + F - 1] = (R)((double)a[((size_t)b * N + i) * F + F - 1] + t);
    }
}

// h_i = node_mask * update_fn(node_mask * [h_i | M_i])      (charge_gn.py:71-74); 64 threads per atom
template <typename R>
__global__ void dense_update_kernel(int n_x, const R* __restrict__ msg, const R* __restrict__ node_mask, UpdW<R> u,
                                    R* __restrict__ a) {
    __shared__ R in[UPD_IN], l1[HID], l2[HID];
    const int F = n_x + HD + 1;
    R* ak = a + (size_t)blockIdx.x * F;
    const R nm = node_mask[blockIdx.x];
    const int t = threadIdx.x;
    if (t < HD) in[t] = ak[n_x + t] * nm;
    if (t < HID) in[HD + t] = msg[(size_t)blockIdx.x * HID + t] * nm;
    __syncthreads();
    if (t < HID) { R s = u.c1[t]; for (int k = 0; k < UPD_IN; ++k) s = fma(in[k], u.U1[k * HID + t], s); l1[t] = relu(s); }
    __syncthreads();
    if (t < HID) { R s = u.c2[t]; for (int k = 0; k < HID; ++k) s = fma(l1[k], u.U2[k * HID + t], s); l2[t] = relu(s); }
    __syncthreads();
    if (t < HD) { R s = u.c3[t]; for (int k = 0; k < HID; ++k) s = fma(l2[k], u.U3[k * HD + t], s); ak[n_x + t] = s * nm; }
}

template <typename R>
__global__ void dense_out_kernel(int n_x, int total, const R* __restrict__ a, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int F = n_x + HD + 1;
    if (i < total) out[i] = (float)a[(size_t)i * F + F - 1];
}

template <typename R>
cudaError_t launch_dense_forward(int B, int N, int n_x, int T, const float* h, const float* e, const float* x, const float* q,
                                 const float* mask, const DenseW<R>* msgw, const UpdW<R>& upd, const DenseW<R>* pasw,
                                 R* a, R* node_mask, R* uv, R* msg, float* q_out, cudaStream_t st) {
    const int rows = B * N;
    dense_untile_kernel<R><<<rows, 64, 0, st>>>(N, n_x, h, x, q, mask, a, node_mask);
    for (int t = 0; t < T; ++t) {
        dense_project_kernel<R><<<rows, 64, 0, st>>>(n_x, a, msgw[t], uv);
        dense_pair_kernel<R, 0><<<rows, DENSE_THREADS, 0, st>>>(N, n_x, e, mask, uv, msgw[t], msg, a);
        dense_update_kernel<R><<<rows, 64, 0, st>>>(n_x, msg, node_mask, upd, a);
    }
    for (int t = 0; t < T; ++t) {
        dense_project_kernel<R><<<rows, 64, 0, st>>>(n_x, a, pasw[t], uv);
        dense_pair_kernel<R, 1><<<rows, DENSE_THREADS, 0, st>>>(N, n_x, e, mask, uv, pasw[t], msg, a);
    }
    dense_out_kernel<R><<<div_up(rows, 256), 256, 0, st>>>(n_x, rows, a, q_out);
    return cudaGetLastError();
}

template cudaError_t launch_dense_forward<float>(int, int, int, int, const float*, const float*, const float*, const float*, const float*,
                                                 const DenseW<float>*, const UpdW<float>&, const DenseW<float>*, float*, float*, float*,
                                                 float*, float*, cudaStream_t);
template cudaError_t launch_dense_forward<double>(int, int, int, int, const float*, const float*, const float*, const float*, const float*,
                                                  const DenseW<double>*, const UpdW<double>&, const DenseW<double>*, double*, double*,
                                                  double*, double*, float*, cudaStream_t);
