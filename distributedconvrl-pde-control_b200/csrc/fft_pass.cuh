// Four-step FFT building block shared by the KS and Navier-Stokes kernels: one pass =
// in-register DFT_P -> transpose through shared memory -> twiddle on the load side -> in-register DFT_Q.
#pragma once
#include "common.cuh"
#include "dft_gen.cuh"

namespace pdeb200 {

template <int P_, int Q_> struct PassStride { static constexpr int value = (Q_ % 2 == 0) ? Q_ + 1 : Q_; };

// Four-step FFT pass.
// in : threads t < Q hold elements (t + Q*r), r < P        (registers zr/zi[0..P))
// out: threads t < P hold elements (t + P*r), r < Q
// transform: X[k] = sum_n x[n] exp(SIGN*2*pi*i*n*k/(P*Q)); tw = [Q][P] forward twiddles W^(n2*k1).
// The twiddle multiply sits on the LOAD side of the transpose so that the twiddle loads are issued
// together with the data loads (one exposed shared-memory latency per pass instead of two).
template <typename T, int P, int Q, int SIGN>
__device__ __forceinline__ void fft_pass(T* __restrict__ zr, T* __restrict__ zi, typename V2<T>::type* xb,
                                         const typename V2<T>::type* __restrict__ tw, int t) {
    using C = typename V2<T>::type;
    constexpr int STRIDE = PassStride<P, Q>::value;
    if (t < Q) {
        dft_r<P, T, SIGN>(zr, zi);
#pragma unroll
        for (int k = 0; k < P; ++k) xb[k * STRIDE + t] = V2<T>::make(zr[k], zi[k]);
    }
    __syncwarp();
    if (t < P) {
        C v[Q], w[Q];
#pragma unroll
        for (int n = 0; n < Q; ++n) v[n] = xb[t * STRIDE + n];
#pragma unroll
        for (int n = 1; n < Q; ++n) w[n] = tw[n * P + t];
        zr[0] = v[0].x; zi[0] = v[0].y;
#pragma unroll
        for (int n = 1; n < Q; ++n) {
            const T wi = (SIGN < 0) ? w[n].y : -w[n].y;
            zr[n] = v[n].x * w[n].x - v[n].y * wi;
            zi[n] = v[n].x * wi + v[n].y * w[n].x;
        }
        dft_r<Q, T, SIGN>(zr, zi);
    }
    __syncwarp();
}

}  // namespace pdeb200
