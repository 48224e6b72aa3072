// KS core kernel, fp64 N = 256 (16 x 16): the many-substep variant with TENSOR MEMORY as a per-thread scratch store.
//
// Same algorithm and layout as ks_step_kernel (ks_step.cuh; KSSetup.jl:130-160): two environments per complex FFT, four-step
// 16 x 16 transform per half warp, u_hat in registers across the substeps.  What changes is where everything else lives.
// ncu on ks_step_kernel (profiles/r1_ks_step_f64.md, r1m) showed the shared-memory crossbar as the higher floor (73 % of its
// wavefront peak against 54 % of the FP64 pipe): per pair and substep 128 wavefronts for the two transposes, but also 64 for the
// N^{n-1} array, 30 for the twiddles and 16 for the coefficient rows -- and the forcing term F held in 64 registers per thread
// to keep another 64 off that crossbar, which pins the kernel at 255 registers.
// Blackwell's tensor memory (256 KB per SM, 128 lanes x 512 32-bit columns) is addressed per lane: tcgen05.ld/st.32x32b give
// every thread of a warp a private row of columns, with its own data path (tools/tmem_probe.cu: > 500 B/clk/SM each way, fully
// overlapped with shared-memory loads).  So every per-thread array of this kernel that is not exchanged between threads lives
// there: N^{n-1} (64 columns), F (64), u_hat itself (64: the CNAB2 update reads the old value and writes the new one, which is
// also the next substep's transform input, so the registers hold ONE 16-point complex line per thread instead of three) and the
// thread's 15 twiddles (64) = 256 columns per thread, two warps per lane quadrant = 512 columns.  Shared memory carries the
// transposes and the two coefficient rows (16 of the former 238 wavefronts per pair and substep).
#pragma once
#include "ks_step.cuh"

namespace pdeb200 {

namespace tm {

#define PDEB_TM_REGS16(r) r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15]

// 16 consecutive 32-bit columns of the calling thread's lane -> 16 registers; asynchronous: wait_ld + touch16 before use
__device__ __forceinline__ void ld16(uint32_t (&r)[16], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void st16(const uint32_t (&r)[16], uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// empty statement that "rewrites" the registers: no use of them can be scheduled above it, and it stays below the preceding
// wait_ld (volatile statements keep their order)
__device__ __forceinline__ void touch16(uint32_t (&r)[16]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ double dbl(const uint32_t (&r)[16], int j) { return __hiloint2double((int)r[2 * j + 1], (int)r[2 * j]); }
__device__ __forceinline__ void put(uint32_t (&r)[16], int j, double v) { r[2 * j] = (uint32_t)__double2loint(v); r[2 * j + 1] = (uint32_t)__double2hiint(v); }

constexpr int kPrev = 0, kF = 64, kU = 128, kTw = 192, kColsPerThread = 256;

// Four-step pass 16 x 16 (fft_pass<double,16,16,SIGN>) with the thread's twiddles coming from tensor memory.
template <int SIGN>
__device__ __forceinline__ void fft_pass16(double* __restrict__ zr, double* __restrict__ zi, double2* xb, uint32_t tbase, int t) {
    constexpr int STRIDE = PassStride<16, 16>::value;
    dft_r<16, double, SIGN>(zr, zi);
#pragma unroll
    for (int k = 0; k < 16; ++k) xb[k * STRIDE + t] = make_double2(zr[k], zi[k]);
    uint32_t w[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) ld16(w[c], tbase + kTw + 16 * c);
    __syncwarp();
    double2 v[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) v[n] = xb[t * STRIDE + n];
    wait_ld();
#pragma unroll
    for (int c = 0; c < 4; ++c) touch16(w[c]);
    zr[0] = v[0].x; zi[0] = v[0].y;
#pragma unroll
    for (int n = 1; n < 16; ++n) {
        const double wr = dbl(w[n / 4], 2 * (n % 4)), wy = dbl(w[n / 4], 2 * (n % 4) + 1);
        const double wi = (SIGN < 0) ? wy : -wy;
        zr[n] = v[n].x * wr - v[n].y * wi;
        zi[n] = v[n].x * wi + v[n].y * wr;
    }
    dft_r<16, double, SIGN>(zr, zi);
    __syncwarp();
}

}  // namespace tm

// Dynamic shared memory: [16 B: TMEM base] [c1 | cN: 2 x 256 doubles] [PAIRS x 16*17 complex: exchange]
//                        [gather path only: sensor table w | idx]
__host__ __device__ inline size_t ks_tm_smem_bytes(int pairs, size_t n_tab) {
    return 16 + 2 * 256 * sizeof(double) + (size_t)pairs * 16 * PassStride<16, 16>::value * sizeof(double2) +
           n_tab * (sizeof(double) + sizeof(int));
}

template <bool SPEC>
__global__ void __launch_bounds__(256, 1) ks_step_tm_kernel(const __grid_constant__ KsArgs<double> A) {
    using namespace tm;
    constexpr int N = 256, R = 16, XB = 16 * PassStride<16, 16>::value;
    const int PAIRS = blockDim.x / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_slot = reinterpret_cast<uint32_t*>(smem_raw);
    double* s_c1 = reinterpret_cast<double*>(smem_raw + 16);
    double* s_cN = s_c1 + N;
    double2* s_xb0 = reinterpret_cast<double2*>(s_cN + N);
    const int warp = threadIdx.x >> 5;
    const int t = threadIdx.x & 15;
    const int pic = threadIdx.x >> 4;
    const int pair = blockIdx.x * PAIRS + pic;
    double2* xb = s_xb0 + (size_t)pic * XB;
    const int n_s = A.n_sensors;
    const size_t n_tab = SPEC ? 0 : (size_t)A.sens.nnz_max * n_s;
    double* s_w = reinterpret_cast<double*>(s_xb0 + (size_t)PAIRS * XB);
    int* s_i = reinterpret_cast<int*>(s_w + n_tab);

    const uint32_t n_cols = blockDim.x > 128 ? 512u : 256u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_slot)), "r"(n_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    const int ea = 2 * pair, eb = 2 * pair + 1;
    const bool va = ea < A.n_envs, vb = eb < A.n_envs;     // pairs beyond the batch run on zeros
    double zr[R], zi[R], ur[R], ui[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int n = t + 16 * r;
        zr[r] = va ? __ldg(A.p + (size_t)ea * N + n) : 0.0;
        zi[r] = vb ? __ldg(A.p + (size_t)eb * N + n) : 0.0;
        ur[r] = va ? A.y[(size_t)ea * N + n] : 0.0;
        ui[r] = vb ? A.y[(size_t)eb * N + n] : 0.0;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) { s_c1[i] = __ldg(A.c1 + i); s_cN[i] = __ldg(A.cN + i); }
    for (size_t i = threadIdx.x; i + 1 <= n_tab; i += blockDim.x) { s_w[i] = __ldg(A.sens.w + i); s_i[i] = __ldg(A.sens.idx + i); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this thread's row of tensor memory: lane = 32 * (warp % 4) + lane id (implicit), columns of warp / 4
    const uint32_t tb = *s_slot + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(warp >> 2) * kColsPerThread;

    // ---- the thread's twiddles W^(n t), n < 16 -> tensor memory ------------------------------------------------------------------
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t q[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 w = __ldg(A.tw12 + (4 * c + j) * 16 + t);
            put(q, 2 * j, w.x); put(q, 2 * j + 1, w.y);
        }
        st16(q, tb + kTw + 16 * c);
    }
    wait_st();

    // ---- job loop (see ks_step_kernel): -2: F = A_inv h fft(p) + h m_hat; -1: u_hat = fft(y); 0..S-1: CNAB2; S: final inverse.
    //      z holds u_hat at the top of every job >= 0: the update writes the new u_hat into z (and into tensor memory) ------------
    for (int job = -2; job <= A.S; ++job) {
        if (job >= 0) {
            fft_pass16<+1>(zr, zi, xb, tb, t);
            if (job == A.S) break;
#pragma unroll
            for (int r = 0; r < R; ++r) { zr[r] = zr[r] * zr[r]; zi[r] = zi[r] * zi[r]; }
        } else if (job == -1) {
#pragma unroll
            for (int r = 0; r < R; ++r) { zr[r] = ur[r]; zi[r] = ui[r]; }
        }
        fft_pass16<-1>(zr, zi, xb, tb, t);
        if (job == -2) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t q[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = 4 * c + j, k = t + 16 * r;
                    const double ah = __ldg(A.ainvh + k);
                    double f_r = ah * zr[r], f_i = ah * zi[r];
                    if (A.hm) { const double2 m = A.hm[k]; f_r += m.x - m.y; f_i += m.x + m.y; }   // m_hat*(1+i)
                    put(q, 2 * j, f_r); put(q, 2 * j + 1, f_i);
                }
                st16(q, tb + kF + 16 * c);
            }
        } else if (job == -1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t q[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) { put(q, 2 * j, zr[4 * c + j]); put(q, 2 * j + 1, zi[4 * c + j]); }
                st16(q, tb + kU + 16 * c);
            }
        } else {
            // u = c1*u + F + i*cN*(N^n - N^{n-1}/3), N^{n-1} := N^n on the first substep (quirk Q2: stored first, then read back
            // like any other substep's).  Rows are (re, im) interleaved per mode -- the register order the transposes' 16-byte
            // stores already want -- and the loads of chunk c + 1 are in flight while chunk c is combined.
            if (job == 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t q[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { put(q, 2 * j, zr[4 * c + j]); put(q, 2 * j + 1, zi[4 * c + j]); }
                    st16(q, tb + kPrev + 16 * c);
                }
                wait_st();
            }
            uint32_t qp[2][16], qf[2][16], qu[2][16];
            ld16(qp[0], tb + kPrev); ld16(qf[0], tb + kF); ld16(qu[0], tb + kU);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int b = c & 1;
                wait_ld();
                touch16(qp[b]); touch16(qf[b]); touch16(qu[b]);
                if (c < 3) {
                    ld16(qp[b ^ 1], tb + kPrev + 16 * (c + 1)); ld16(qf[b ^ 1], tb + kF + 16 * (c + 1)); ld16(qu[b ^ 1], tb + kU + 16 * (c + 1));
                }
                uint32_t qs[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) { put(qs, 2 * j, zr[4 * c + j]); put(qs, 2 * j + 1, zi[4 * c + j]); }
                st16(qs, tb + kPrev + 16 * c);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = 4 * c + j, k = t + 16 * r;
                    const double tr = fma(-A.third, dbl(qp[b], 2 * j), zr[r]);
                    const double ti = fma(-A.third, dbl(qp[b], 2 * j + 1), zi[r]);
                    const double c1 = s_c1[k], cn = s_cN[k];
                    zr[r] = fma(-cn, ti, fma(c1, dbl(qu[b], 2 * j), dbl(qf[b], 2 * j)));
                    zi[r] = fma(cn, tr, fma(c1, dbl(qu[b], 2 * j + 1), dbl(qf[b], 2 * j + 1)));
                }
                uint32_t qn[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) { put(qn, 2 * j, zr[4 * c + j]); put(qn, 2 * j + 1, zi[4 * c + j]); }
                st16(qn, tb + kU + 16 * c);
            }
        }
        wait_st();
    }

    // ---- y = ifft(u_hat) / N: store, max |y| (PDEenv.jl:227), sensor dots (KSSetup.jl:168-170) ------------------------------------
    double vmax_a = 0.0, vmax_b = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int n = t + 16 * r;
        const double ya = zr[r] * A.inv_n, yb = zi[r] * A.inv_n;
        if (!SPEC) xb[A.perm ? __ldg(A.perm + n) : n] = make_double2(ya, yb);
        if (va) A.y[(size_t)ea * N + n] = ya;
        if (vb) A.y[(size_t)eb * N + n] = yb;
        vmax_a = fmax(vmax_a, fabs(ya)); vmax_b = fmax(vmax_b, fabs(yb));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        vmax_a = fmax(vmax_a, __shfl_xor_sync(0xffffffffu, vmax_a, o, 16));
        vmax_b = fmax(vmax_b, __shfl_xor_sync(0xffffffffu, vmax_b, o, 16));
    }
    if (t == 0) {
        if (va) A.vmax_out[ea] = vmax_a;
        if (vb) A.vmax_out[eb] = vmax_b;
    }
    if (SPEC) {
        // one more inverse transform of u_hat . conj(fft(g_0)): <y, g_i> = result[sens_sp * i]   (ks.cu::ks_bases_changed)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t qu[16];
            ld16(qu, tb + kU + 16 * c);
            wait_ld();
            touch16(qu);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = 4 * c + j;
                const double2 h = __ldg(A.sens_hat + t + 16 * r);
                const double u_r = dbl(qu, 2 * j), u_i = dbl(qu, 2 * j + 1);
                zr[r] = u_r * h.x - u_i * h.y;
                zi[r] = u_r * h.y + u_i * h.x;
            }
        }
        fft_pass16<+1>(zr, zi, xb, tb, t);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = t + 16 * r;
            const int i = n / A.sens_sp;
            if (i * A.sens_sp == n && i < n_s) {
                if (va) A.sensors_out[(size_t)ea * n_s + i] = zr[r] * A.inv_n;
                if (vb) A.sensors_out[(size_t)eb * n_s + i] = zi[r] * A.inv_n;
            }
        }
    } else {
        __syncwarp();
        constexpr int CU = 4;
        for (int i0 = t; i0 < n_s; i0 += CU * 16) {
            double sa[CU], sb[CU];
#pragma unroll
            for (int m = 0; m < CU; ++m) { sa[m] = 0.0; sb[m] = 0.0; }
#pragma unroll 3
            for (int j = 0; j < A.sens.nnz_max; ++j) {
#pragma unroll
                for (int m = 0; m < CU; ++m) {
                    const int i = i0 + m * 16;
                    const int idx = i < n_s ? s_i[j * n_s + i] : 0;
                    const double w = i < n_s ? s_w[j * n_s + i] : 0.0;
                    const double2 v = xb[idx];
                    sa[m] += v.x * w; sb[m] += v.y * w;
                }
            }
#pragma unroll
            for (int m = 0; m < CU; ++m) {
                const int i = i0 + m * 16;
                if (i < n_s) {
                    if (va) A.sensors_out[(size_t)ea * n_s + i] = sa[m];
                    if (vb) A.sensors_out[(size_t)eb * n_s + i] = sb[m];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*s_slot), "r"(n_cols) : "memory");
}

}  // namespace pdeb200
