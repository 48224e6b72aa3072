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
//   fft_batch_nt : natural in,    transposed out   (column pass DFT_P2, twiddle, row pass DFT_P1)
//   fft_batch_tn : transposed in, natural out      (row pass DFT_P1, twiddle, column pass DFT_P2)
// so an inverse/forward pair (and the point-wise work between them) never needs a reorder.
//
// Twiddles: the column tasks of both forms sit on lane % P1 == column whenever P1 divides 32, so the P2 - 1 factors
// W^(column * k) a lane ever needs are the same for every round, line, job and direction (conjugated for the other
// sign): BatchTwiddles keeps them in registers (TWREG) -- no table traffic on the shared-memory pipe, which is what
// bounds these kernels.  P1 = 12 (192 points) reads the fft_pass tables from shared memory instead.
//
// The first pass of either form takes its input through a functor, so a caller can expand / combine its raw data
// while loading (kernel B: conjugate extension of the half spectrum; products of two lines).
#pragma once
#include "common.cuh"
#include "dft_gen.cuh"

namespace pdeb200 {

template <int P1, int P2> struct BatchLayout {
    static constexpr int N = P1 * P2;
    static constexpr int S = P1 + 1;
    static constexpr int LS = P2 * S;                     // entries per line
    static constexpr bool TWREG = (32 % P1) == 0;
    __host__ __device__ static constexpr int nat(int e) { return (e / P1) * S + e % P1; }
    __host__ __device__ static constexpr int tra(int e) { return (e % P2) * S + e / P2; }
};

// W^(c k) = exp(-2 pi i c k / N), c = lane % P1, k < P2 (forward sign), read once from the table of
// fft_pass<T, P2, P1, .> (tw[n*P2 + k] = W^(n k)).
template <typename T, int P1, int P2> struct BatchTwiddles {
    T re[P2], im[P2];
    __device__ __forceinline__ void load(const typename V2<T>::type* __restrict__ tw_p2p1, int lane) {
        const int c = lane % P1;
#pragma unroll
        for (int k = 0; k < P2; ++k) { const typename V2<T>::type v = tw_p2p1[c * P2 + k]; re[k] = v.x; im[k] = v.y; }
    }
};

struct BatchNoSync { static constexpr bool value = false; };
struct BatchSync { static constexpr bool value = true; };

// X[k] = sum_n x[n] exp(SIGN 2 pi i n k / N).
// col_load(line, column, r) -> element column + P1*r of that line (r is a compile-time-unrolled index).
// SYNC: the loader reads slots other tasks of the same round will overwrite -> __syncwarp between loads and stores
// (a round always covers whole lines when P1 divides 32).
// tw_smem: table of fft_pass<T, P2, P1, .> (tw[n*P2 + k1]); only read when !TWREG.
// The column-pass rounds run from the LAST lines to the first; round_hook(round) is called by every lane at the top
// of a round (kernel B waits there for the bulk copies of that round's lines, the first lines arriving last).
template <typename T, int P1, int P2, int L, int SIGN, bool TWREG, typename SyncT, typename ColLoad, typename RoundHook>
__device__ __forceinline__ void fft_batch_nt(typename V2<T>::type* xb, const typename V2<T>::type* __restrict__ tw_smem,
                                             const BatchTwiddles<T, P1, P2>& tw, int lane, ColLoad col_load,
                                             RoundHook round_hook) {
    using C = typename V2<T>::type;
    using G = BatchLayout<P1, P2>;
    static_assert(!SyncT::value || (L * P1) % 32 == 0, "in-place loaders need whole rounds");
#pragma unroll 1
    for (int id = lane + 32 * ((L * P1 - 1) / 32); id >= 0; id -= 32) {
        round_hook(id >> 5);
        if ((L * P1) % 32 != 0 && id >= L * P1) continue;
        const int line = id / P1, col = id % P1;
        C* b = xb + line * G::LS + col;
        T zr[P2], zi[P2];
#pragma unroll
        for (int r = 0; r < P2; ++r) { const C v = col_load(line, col, r); zr[r] = v.x; zi[r] = v.y; }
        if (SyncT::value) __syncwarp();
        dft_r<P2, T, SIGN>(zr, zi);
        b[0] = V2<T>::make(zr[0], zi[0]);
#pragma unroll
        for (int k = 1; k < P2; ++k) {
            if (TWREG) {
                const T wi = (SIGN < 0) ? tw.im[k] : -tw.im[k];
                b[k * G::S] = V2<T>::make(zr[k] * tw.re[k] - zi[k] * wi, zr[k] * wi + zi[k] * tw.re[k]);
            } else {
                b[k * G::S] = V2<T>::make(zr[k], zi[k]);
            }
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int id = lane; id < L * P2; id += 32) {
        const int k1 = id % P2;
        C* b = xb + (id / P2) * G::LS + k1 * G::S;
        T zr[P1], zi[P1];
        if (TWREG) {
#pragma unroll
            for (int n = 0; n < P1; ++n) { const C v = b[n]; zr[n] = v.x; zi[n] = v.y; }
        } else {
            C v[P1], w[P1];
#pragma unroll
            for (int n = 0; n < P1; ++n) v[n] = b[n];
#pragma unroll
            for (int n = 1; n < P1; ++n) w[n] = tw_smem[n * P2 + k1];
            zr[0] = v[0].x; zi[0] = v[0].y;
#pragma unroll
            for (int n = 1; n < P1; ++n) {
                const T wi = (SIGN < 0) ? w[n].y : -w[n].y;
                zr[n] = v[n].x * w[n].x - v[n].y * wi;
                zi[n] = v[n].x * wi + v[n].y * w[n].x;
            }
        }
        dft_r<P1, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P1; ++k) b[k] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
}

// row_load(line, row, r) -> element row + P2*r of that line (transposed layout: slot row*S + r).
// tw_smem: table of fft_pass<T, P1, P2, .> (tw[t*P1 + k1]); only read when !TWREG.
// mid() is called by every lane between the passes (after the __syncwarp).
template <typename T, int P1, int P2, int L, int SIGN, bool TWREG, typename RowLoad, typename Mid>
__device__ __forceinline__ void fft_batch_tn(typename V2<T>::type* xb, const typename V2<T>::type* __restrict__ tw_smem,
                                             const BatchTwiddles<T, P1, P2>& tw, int lane, RowLoad row_load, Mid mid) {
    using C = typename V2<T>::type;
    using G = BatchLayout<P1, P2>;
#pragma unroll 1
    for (int id = lane; id < L * P2; id += 32) {
        const int line = id / P2, row = id % P2;
        C* b = xb + line * G::LS + row * G::S;
        T zr[P1], zi[P1];
#pragma unroll
        for (int r = 0; r < P1; ++r) { const C v = row_load(line, row, r); zr[r] = v.x; zi[r] = v.y; }
        dft_r<P1, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P1; ++k) b[k] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
    mid();
#pragma unroll 1
    for (int id = lane; id < L * P1; id += 32) {
        const int k1 = id % P1;
        C* b = xb + (id / P1) * G::LS + k1;
        T zr[P2], zi[P2];
        {
            const C v0 = b[0];
            zr[0] = v0.x; zi[0] = v0.y;
        }
#pragma unroll
        for (int t = 1; t < P2; ++t) {
            const C v = b[t * G::S];
            T wr, wi;
            if (TWREG) { wr = tw.re[t]; wi = (SIGN < 0) ? tw.im[t] : -tw.im[t]; }
            else { const C w = tw_smem[t * P1 + k1]; wr = w.x; wi = (SIGN < 0) ? w.y : -w.y; }
            zr[t] = v.x * wr - v.y * wi;
            zi[t] = v.x * wi + v.y * wr;
        }
        dft_r<P2, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P2; ++k) b[k * G::S] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
}

}  // namespace pdeb200
