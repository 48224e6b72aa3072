// Batched in-place four-step FFT: one warp transforms L lines of N = P1*P2 points that sit in shared memory.
//
// fft_pass (fft_pass.cuh) keeps one line per warp in registers: DFT_P on Q lanes, DFT_Q on P lanes -- for
// 384 = 16 x 24 that is 16 resp. 24 of 32 lanes per FP64 instruction.  Here the (line, sub-transform) pairs of L
// lines are dealt to the lanes as one task list, so every round of a pass has 32 busy lanes (L = 4, 384 points:
// 64 DFT_24 tasks = 2 rounds, 96 DFT_16 tasks = 3 rounds; 37 % fewer FP64 warp instructions per line).
//
// A line is a P2 x P1 matrix with row stride S = P1 + 1 (odd: the row pass is bank-conflict free):
//   natural layout     element e at (e / P1) * S + e % P1
//   transposed layout  element e at (e % P2) * S + e / P2
// Both passes are in place PER TASK (a task reads one column resp. one row and writes the same column / row), so
// the only synchronisation is one __syncwarp between the passes:
//   fft_batch_nt : natural in,    transposed out   (column pass DFT_P2, then twiddle + row pass DFT_P1)
//   fft_batch_tn : transposed in, natural out      (row pass DFT_P1, then twiddle + column pass DFT_P2)
// so an inverse/forward pair (and the point-wise work between them) never needs a reorder.
#pragma once
#include "common.cuh"
#include "dft_gen.cuh"

namespace pdeb200 {

template <int P1, int P2> struct BatchLayout {
    static constexpr int N = P1 * P2;
    static constexpr int S = P1 + 1;
    static constexpr int LS = P2 * S;                     // entries per line
    __host__ __device__ static constexpr int nat(int e) { return (e / P1) * S + e % P1; }
    __host__ __device__ static constexpr int tra(int e) { return (e % P2) * S + e / P2; }
};

// X[k] = sum_n x[n] exp(SIGN 2 pi i n k / N).  tw = the table of fft_pass<T, P2, P1, .>: tw[n*P2 + k1] = W^(n k1).
template <typename T, int P1, int P2, int L, int SIGN>
__device__ __forceinline__ void fft_batch_nt(typename V2<T>::type* xb, const typename V2<T>::type* __restrict__ tw, int lane) {
    using C = typename V2<T>::type;
    using G = BatchLayout<P1, P2>;
#pragma unroll 1
    for (int id = lane; id < L * P1; id += 32) {
        C* b = xb + (id / P1) * G::LS + id % P1;
        T zr[P2], zi[P2];
#pragma unroll
        for (int r = 0; r < P2; ++r) { const C v = b[r * G::S]; zr[r] = v.x; zi[r] = v.y; }
        dft_r<P2, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P2; ++k) b[k * G::S] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
#pragma unroll 1
    for (int id = lane; id < L * P2; id += 32) {
        const int k1 = id % P2;
        C* b = xb + (id / P2) * G::LS + k1 * G::S;
        C v[P1], w[P1];
#pragma unroll
        for (int n = 0; n < P1; ++n) v[n] = b[n];
#pragma unroll
        for (int n = 1; n < P1; ++n) w[n] = tw[n * P2 + k1];
        T zr[P1], zi[P1];
        zr[0] = v[0].x; zi[0] = v[0].y;
#pragma unroll
        for (int n = 1; n < P1; ++n) {
            const T wi = (SIGN < 0) ? w[n].y : -w[n].y;
            zr[n] = v[n].x * w[n].x - v[n].y * wi;
            zi[n] = v[n].x * wi + v[n].y * w[n].x;
        }
        dft_r<P1, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P1; ++k) b[k] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
}

// tw = the table of fft_pass<T, P1, P2, .>: tw[t*P1 + k1] = W^(t k1), t < P2, k1 < P1.
template <typename T, int P1, int P2, int L, int SIGN>
__device__ __forceinline__ void fft_batch_tn(typename V2<T>::type* xb, const typename V2<T>::type* __restrict__ tw, int lane) {
    using C = typename V2<T>::type;
    using G = BatchLayout<P1, P2>;
#pragma unroll 1
    for (int id = lane; id < L * P2; id += 32) {
        C* b = xb + (id / P2) * G::LS + (id % P2) * G::S;
        T zr[P1], zi[P1];
#pragma unroll
        for (int r = 0; r < P1; ++r) { const C v = b[r]; zr[r] = v.x; zi[r] = v.y; }
        dft_r<P1, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P1; ++k) b[k] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
#pragma unroll 1
    for (int id = lane; id < L * P1; id += 32) {
        const int k1 = id % P1;
        C* b = xb + (id / P1) * G::LS + k1;
        C v[P2], w[P2];
#pragma unroll
        for (int t = 0; t < P2; ++t) v[t] = b[t * G::S];
#pragma unroll
        for (int t = 1; t < P2; ++t) w[t] = tw[t * P1 + k1];
        T zr[P2], zi[P2];
        zr[0] = v[0].x; zi[0] = v[0].y;
#pragma unroll
        for (int t = 1; t < P2; ++t) {
            const T wi = (SIGN < 0) ? w[t].y : -w[t].y;
            zr[t] = v[t].x * w[t].x - v[t].y * wi;
            zi[t] = v[t].x * wi + v[t].y * w[t].x;
        }
        dft_r<P2, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P2; ++k) b[k * G::S] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
}

}  // namespace pdeb200
