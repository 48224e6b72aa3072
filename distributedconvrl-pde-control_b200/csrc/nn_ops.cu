// Network operator entry points: the callable CustomNeuralNetworkApproximator of the reference
// (src/custom_nna.jl:13, `(app)(x) = app.model(x)` on a (rows, n_columns) matrix) for arbitrary column counts,
// layer by layer:  dense layers (in, out >= 32) on the tensor cores (dense_tc.cuh, 3xTF32 tcgen05), thin ones
// (the K <= 13 input layers and the 1-wide output heads of the shipped networks) on CUDA cores.
#include <algorithm>
#include <string>
#include <vector>

#include "ctx.hpp"
#include "dense_tc.cuh"

namespace pdeb200 {
namespace {

inline int pad4(int n) { return (n + 3) / 4 * 4; }

// Weight operand for the tensor-core kernels, K-major rows with leading dimension ld, split into TF32 hi + remainder lo:
//   transpose = 1: out[n][k] = W[n + N*k]  (forward:  Flux (out, in) column-major -> rows n, contraction k)
//   transpose = 0: out[k][n] = W[n + N*k]  (dgrad:    the Flux memory itself, rows k, contraction n; only re-pitched)
__global__ void split_w_kernel(int N, int K, int ld, int transpose, const float* __restrict__ W, float* __restrict__ hi,
                               float* __restrict__ lo) {
    const int rows = transpose ? N : K, cols = transpose ? K : N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld) return;
    const int r = i / ld, c = i % ld;
    float w = 0.f;
    if (c < cols) w = transpose ? W[r + (size_t)N * c] : W[c + (size_t)N * r];
    const float h = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    hi[i] = h; lo[i] = w - h;
}

// CUDA-core Dense for thin contractions: one thread per (column, output unit), x row in registers via L1.
__global__ void dense_thin_kernel(int M, int K, int N, const float* __restrict__ X, long long ldx, const float* __restrict__ W,
                                  const float* __restrict__ b, int act, float* __restrict__ Y, long long ldy) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= (long long)M * N) return;
    const int m = (int)(q / N), n = (int)(q % N);
    const float* x = X + (long long)m * ldx;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(__ldg(W + n + (size_t)N * k), x[k], acc);
    Y[(long long)m * ldy + n] = tc::act_f(act, acc + b[n]);
}

// Thin contraction, wide output (e.g. 13 -> 340): four output units per thread, float4 weight loads and stores.
__global__ void dense_thin4_kernel(int M, int K, int N, const float* __restrict__ X, long long ldx, const float* __restrict__ W,
                                   const float* __restrict__ b, int act, float* __restrict__ Y, long long ldy) {
    const int n4 = N >> 2;
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= (long long)M * n4) return;
    const int m = (int)(q / n4), n = (int)(q % n4) * 4;
    const float* x = X + (long long)m * ldx;
    float4 acc = __ldg(reinterpret_cast<const float4*>(b + n));
    for (int k = 0; k < K; ++k) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + n + (size_t)N * k));
        const float xv = x[k];
        acc.x = fmaf(w.x, xv, acc.x); acc.y = fmaf(w.y, xv, acc.y); acc.z = fmaf(w.z, xv, acc.z); acc.w = fmaf(w.w, xv, acc.w);
    }
    acc.x = tc::act_f(act, acc.x); acc.y = tc::act_f(act, acc.y); acc.z = tc::act_f(act, acc.z); acc.w = tc::act_f(act, acc.w);
    *reinterpret_cast<float4*>(Y + (long long)m * ldy + n) = acc;
}

// Narrow heads (N <= 8, e.g. the 340 -> 1 output of the critic): one warp per column, lanes stride over K so the
// row is read coalesced, shuffle reduction.
template <int NMAX>
__global__ void dense_head_kernel(int M, int K, int N, const float* __restrict__ X, long long ldx, const float* __restrict__ W,
                                  const float* __restrict__ b, int act, float* __restrict__ Y, long long ldy) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* x = X + (long long)m * ldx;
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float xv = x[k];
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < N) acc[n] = fmaf(__ldg(W + n + (size_t)N * k), xv, acc[n]);
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    if (lane == 0)
        for (int n = 0; n < N; ++n) Y[(long long)m * ldy + n] = tc::act_f(act, acc[n] + b[n]);
}

__global__ void pad_rows_kernel(long long M, int K, int ld, float* X) {       // zero the padding columns [K, ld)
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int w = ld - K;
    if (w <= 0 || q >= M * w) return;
    X[(q / w) * ld + K + (q % w)] = 0.f;
}

struct Scratch {
    float* buf[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    float* wt = nullptr; size_t wt_cap = 0;
};
thread_local Scratch g_scr;

// cuTensorMapEncodeTiled through the runtime (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 2-D fp32 tensor map over a row-major [rows][ld] array: box = BK x 128 (128 bytes x 128 rows), SWIZZLE_128B, zero fill
int32_t make_weight_map(pdeb200_ctx* c, CUtensorMap* map, float* ptr, int rows, int cols, int ld) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return fail(c, PDEB200_ECUDA, "dense: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)tc::BN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, PDEB200_ECUDA, "dense: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return PDEB200_OK;
}

int sm_count(const pdeb200_ctx* c) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, c->device);
    return n;
}

template <bool T, bool B>
int32_t configure_tc(pdeb200_ctx* c) {
    PDEB_CUDA(c, ensure_dyn_smem(tc::dense_tc_kernel<T, B>, (size_t)tc::SMEM_BYTES, c->device));
    return PDEB200_OK;
}

int32_t ensure(pdeb200_ctx* c, float** p, size_t* cap, size_t need) {
    if (need <= *cap) return PDEB200_OK;
    if (*p) cudaFree(*p);
    *cap = 0; *p = nullptr;
    PDEB_CUDA(c, cudaMalloc(p, need * sizeof(float)));
    *cap = need;
    return PDEB200_OK;
}

}  // namespace

// one Dense layer on device buffers; path: 0 auto, 1 CUDA cores, 2 tensor cores
int32_t dense_layer(pdeb200_ctx* c, int M, int K, int N, const float* X, long long ldx, const float* W, int act, float* Y,
                    long long ldy, int path, int* used_tc) {
    const float* bias = W + (size_t)K * N;
    const bool can_tc = (ldx % 4 == 0) && K >= 8;
    const bool want_tc = path == 2 || (path == 0 && K >= 32 && N >= 32 && M >= 64);
    if (path == 2 && !can_tc) return fail(c, PDEB200_EUNSUPPORTED, "dense: tensor-core path needs K >= 8 and a 16-byte aligned leading dimension");
    if (want_tc && can_tc) {
        const int ldw = pad4(K);
        int32_t rc = ensure(c, &g_scr.wt, &g_scr.wt_cap, (size_t)2 * N * ldw);
        if (rc) return rc;
        float* whi = g_scr.wt; float* wlo = g_scr.wt + (size_t)N * ldw;
        split_w_kernel<<<(N * ldw + 255) / 256, 256, 0, c->stream>>>(N, K, ldw, 1, W, whi, wlo);
        tc::DenseTmaMaps TM;
        if ((rc = make_weight_map(c, &TM.w_hi, whi, N, ldw, ldw)) || (rc = make_weight_map(c, &TM.w_lo, wlo, N, ldw, ldw))) return rc;
        tc::DenseArgs A;
        A.X = X; A.ldx = ldx; A.Wt = whi; A.ldw = ldw; A.bias = bias; A.Y = Y; A.ldy = ldy;
        A.M = M; A.N = N; A.K = ldw; A.act = act;      // padded K columns are zero in the weights; X's are zeroed by the caller
        A.mask = nullptr; A.ldm = 0; A.mask_act = 0; A.split_len = 0; A.y_split_stride = 0;
        if ((rc = configure_tc<false, true>(c))) return rc;
        const int n_tiles = ((M + tc::BM - 1) / tc::BM) * ((N + tc::BN - 1) / tc::BN);
        tc::dense_tc_kernel<false, true><<<std::min(n_tiles, sm_count(c)), tc::N_THREADS, tc::SMEM_BYTES, c->stream>>>(A, TM);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 2;
        if (used_tc) *used_tc += 1;
        return PDEB200_OK;
    }
    const long long total = (long long)M * N;
    if (N <= 8 && K >= 64) {
        if (N == 1) dense_head_kernel<1><<<(M + 7) / 8, 256, 0, c->stream>>>(M, K, N, X, ldx, W, bias, act, Y, ldy);
        else dense_head_kernel<8><<<(M + 7) / 8, 256, 0, c->stream>>>(M, K, N, X, ldx, W, bias, act, Y, ldy);
    } else if (N % 4 == 0 && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
               ((size_t)K * N) % 4 == 0) {
        const long long t4 = total / 4;
        dense_thin4_kernel<<<(unsigned)((t4 + 255) / 256), 256, 0, c->stream>>>(M, K, N, X, ldx, W, bias, act, Y, ldy);
    } else
        dense_thin_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(M, K, N, X, ldx, W, bias, act, Y, ldy);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

namespace {

__global__ void dgrad_thin_kernel(int M, int K, int N, const float* __restrict__ dY, long long lddy, const float* __restrict__ W,
                                  const float* __restrict__ mask, long long ldm, int mask_act, float* __restrict__ dX, long long lddx) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= (long long)M * K) return;
    const int m = (int)(q / K), k = (int)(q % K);
    const float* d = dY + (long long)m * lddy;
    const float* w = W + (size_t)N * k;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc = fmaf(__ldg(w + n), d[n], acc);
    if (mask) acc *= tc::act_grad_f(mask_act, mask[(long long)m * ldm + k]);
    dX[(long long)m * lddx + k] = acc;
}

// partial[z][n + N*k] = sum over the z-th slice of m of dY[m][n] * X[m][k]   (thin layers: N*K small)
__global__ void wgrad_thin_kernel(int M, int K, int N, int slice, const float* __restrict__ dY, long long lddy,
                                  const float* __restrict__ X, long long ldx, float* __restrict__ partial) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N * K) return;
    const int n = q % N, k = q / N;
    const int m0 = blockIdx.y * slice, m1 = min(M, m0 + slice);
    float acc = 0.f;
    int m = m0;
    for (; m + 8 <= m1; m += 8) {             // loads first, FMAs after, ascending m
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { a[u] = __ldg(dY + (long long)(m + u) * lddy + n); b[u] = __ldg(X + (long long)(m + u) * ldx + k); }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = fmaf(a[u], b[u], acc);
    }
    for (; m < m1; ++m) acc = fmaf(__ldg(dY + (long long)m * lddy + n), __ldg(X + (long long)m * ldx + k), acc);
    partial[(size_t)blockIdx.y * N * K + q] = acc;
}

// partial[z][n] = sum over the z-th slice of m of dY[m][n]
__global__ void colsum_kernel(int M, int N, int slice, const float* __restrict__ dY, long long lddy, float* __restrict__ partial) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int m0 = blockIdx.y * slice, m1 = min(M, m0 + slice);
    float acc = 0.f;
    int m = m0;
    for (; m + 8 <= m1; m += 8) {
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = __ldg(dY + (long long)(m + u) * lddy + n);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += a[u];
    }
    for (; m < m1; ++m) acc += __ldg(dY + (long long)m * lddy + n);
    partial[(size_t)blockIdx.y * N + n] = acc;
}

// out[n + N*k] (transpose = 1: partial tiles are [n][ldp] row-major) or out[q] = sum_z partial[z][...], fixed order
__global__ void reduce_splits_kernel(int Z, int N, int K, int ldp, int transposed, const float* __restrict__ partial,
                                     long long zstride, float* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N * K) return;
    const int n = q % N, k = q / N;
    const long long src = transposed ? (long long)n * ldp + k : q;
    double s = 0.0;
    int z = 0;
    for (; z + 8 <= Z; z += 8) {              // loads first, adds after (in-order issue), ascending z (fixed order)
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(partial + (long long)(z + u) * zstride + src);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; z < Z; ++z) s += (double)__ldg(partial + (long long)z * zstride + src);
    out[q] = (float)s;
}

struct GradScratch { float* p = nullptr; size_t cap = 0; float* wp = nullptr; size_t wp_cap = 0; };
thread_local GradScratch g_gs;

}  // namespace

// dX[M x K] = (dY[M x N] * W) (.*) act'(mask)   -- backward through Dense to its input (Zygote pullback of W*x)
int32_t dense_dgrad(pdeb200_ctx* c, int M, int K, int N, const float* dY, long long lddy, const float* W, const float* mask,
                    long long ldm, int mask_act, float* dX, long long lddx, int path) {
    const bool can_tc = lddy % 4 == 0 && N >= 8;
    const bool want_tc = path == 2 || (path == 0 && K >= 32 && N >= 32 && M >= 64);
    if (want_tc && can_tc) {
        const int ldp = pad4(N);
        int32_t rc = ensure(c, &g_gs.wp, &g_gs.wp_cap, (size_t)2 * K * ldp);
        if (rc) return rc;
        float* whi = g_gs.wp; float* wlo = g_gs.wp + (size_t)K * ldp;
        split_w_kernel<<<(K * ldp + 255) / 256, 256, 0, c->stream>>>(N, K, ldp, 0, W, whi, wlo);
        tc::DenseTmaMaps TM;
        if ((rc = make_weight_map(c, &TM.w_hi, whi, K, ldp, ldp)) || (rc = make_weight_map(c, &TM.w_lo, wlo, K, ldp, ldp))) return rc;
        tc::DenseArgs A;
        A.X = dY; A.ldx = lddy; A.Wt = whi; A.ldw = ldp; A.bias = nullptr; A.Y = dX; A.ldy = lddx;
        A.M = M; A.N = K; A.K = ldp; A.act = 0; A.mask = mask; A.ldm = ldm; A.mask_act = mask_act;
        A.split_len = 0; A.y_split_stride = 0;
        if ((rc = configure_tc<false, true>(c))) return rc;
        const int n_tiles = ((M + tc::BM - 1) / tc::BM) * ((K + tc::BN - 1) / tc::BN);
        tc::dense_tc_kernel<false, true><<<std::min(n_tiles, sm_count(c)), tc::N_THREADS, tc::SMEM_BYTES, c->stream>>>(A, TM);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 2;
        return PDEB200_OK;
    }
    const long long total = (long long)M * K;
    dgrad_thin_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(M, K, N, dY, lddy, W, mask, ldm, mask_act, dX, lddx);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

// gW[n + N*k] = sum_m dY[m][n] X[m][k],  gb[n] = sum_m dY[m][n]   (flat Flux order: W column-major (out,in), then b)
int32_t dense_wgrad(pdeb200_ctx* c, int M, int K, int N, const float* dY, long long lddy, const float* X, long long ldx,
                    float* gW, float* gb, int path) {
    const bool want_tc = path == 2 || (path == 0 && K >= 32 && N >= 32 && M >= 256);
    int32_t rc;
    {   // one scratch allocation for every partial buffer used below (growing it later would sync the device)
        const size_t zmax = (size_t)std::max((M + 255) / 256, 128);
        if ((rc = ensure(c, &g_gs.p, &g_gs.cap, zmax * (size_t)N * pad4(K)))) return rc;
    }
    if (want_tc) {
        const int ldp = pad4(K);
        const int tiles = ((N + tc::BM - 1) / tc::BM) * ((K + tc::BN - 1) / tc::BN);
        int Z = std::max(1, std::min((M + 255) / 256, (2 * 148 + tiles - 1) / tiles));
        const int split_len = ((M + Z - 1) / Z + tc::BK - 1) / tc::BK * tc::BK;
        Z = (M + split_len - 1) / split_len;
        const long long zstride = (long long)N * ldp;
        if ((rc = ensure(c, &g_gs.p, &g_gs.cap, (size_t)Z * zstride))) return rc;
        tc::DenseArgs A;
        A.X = dY; A.ldx = lddy; A.Wt = X; A.ldw = ldx; A.bias = nullptr; A.Y = g_gs.p; A.ldy = ldp;
        A.M = N; A.N = K; A.K = M; A.act = 0; A.mask = nullptr; A.ldm = 0; A.mask_act = 0;
        A.split_len = split_len; A.y_split_stride = zstride;
        if ((rc = configure_tc<true, false>(c))) return rc;
        const int n_tiles = ((N + tc::BM - 1) / tc::BM) * ((K + tc::BN - 1) / tc::BN) * Z;
        tc::DenseTmaMaps TM{};
        tc::dense_tc_kernel<true, false><<<std::min(n_tiles, sm_count(c)), tc::N_THREADS, tc::SMEM_BYTES, c->stream>>>(A, TM);
        reduce_splits_kernel<<<(N * K + 255) / 256, 256, 0, c->stream>>>(Z, N, K, ldp, 1, g_gs.p, zstride, gW);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 2;
    } else {
        const int Z = std::max(1, std::min(128, M / 64));
        const int slice = (M + Z - 1) / Z;
        if ((rc = ensure(c, &g_gs.p, &g_gs.cap, (size_t)Z * N * K))) return rc;
        wgrad_thin_kernel<<<dim3((N * K + 127) / 128, Z), 128, 0, c->stream>>>(M, K, N, slice, dY, lddy, X, ldx, g_gs.p);
        reduce_splits_kernel<<<(N * K + 255) / 256, 256, 0, c->stream>>>(Z, N, K, 0, 0, g_gs.p, (long long)N * K, gW);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 2;
    }
    {
        const int Z = std::max(1, std::min(128, M / 64));
        const int slice = (M + Z - 1) / Z;
        if ((rc = ensure(c, &g_gs.p, &g_gs.cap, (size_t)Z * N))) return rc;
        colsum_kernel<<<dim3((N + 127) / 128, Z), 128, 0, c->stream>>>(M, N, slice, dY, lddy, g_gs.p);
        reduce_splits_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(Z, N, 1, 0, 0, g_gs.p, (long long)N, gb);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 2;
    }
    return PDEB200_OK;
}

}  // namespace pdeb200

using namespace pdeb200;

extern "C" {

int32_t pdeb200_net_forward_device(pdeb200_ctx* c, int32_t net, int32_t n_cols, const float* x_dev, int64_t ldx, float* y_dev,
                                   int64_t ldy, int32_t path, int32_t* n_tensor_layers) {
    if (!c || net < 0 || net > 3 || n_cols < 1 || !x_dev || !y_dev) return fail(c, PDEB200_EINVAL, "net_forward: bad argument");
    const HostNet& n = c->nets[net];
    if (!n.n_layers) return fail(c, PDEB200_ESTATE, "net_forward: network not set");
    if (ldx < n.sizes[0] || ldy < n.sizes[n.n_layers]) return fail(c, PDEB200_EINVAL, "net_forward: leading dimension too small");
    cudaSetDevice(c->device);
    int used = 0;
    const float* in = x_dev; long long ldin = ldx;
    for (int l = 0; l < n.n_layers; ++l) {
        const bool last = l == n.n_layers - 1;
        float* out; long long ldout;
        if (last) { out = y_dev; ldout = ldy; }
        else {
            ldout = pad4(n.sizes[l + 1]);
            int32_t rc = ensure(c, &g_scr.buf[l & 1], &g_scr.cap[l & 1], (size_t)n_cols * ldout);
            if (rc) return rc;
            out = g_scr.buf[l & 1];
            if (ldout != n.sizes[l + 1]) {
                const long long tot = (long long)n_cols * (ldout - n.sizes[l + 1]);
                pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(n_cols, n.sizes[l + 1], (int)ldout, out);
            }
        }
        int32_t rc = dense_layer(c, n_cols, n.sizes[l], n.sizes[l + 1], in, ldin, n.d_params + n.offs[l], n.acts[l], out, ldout,
                                 path, &used);
        if (rc) return rc;
        in = out; ldin = ldout;
    }
    if (n_tensor_layers) *n_tensor_layers = used;
    return PDEB200_OK;
}

int32_t pdeb200_net_forward(pdeb200_ctx* c, int32_t net, int32_t n_cols, const float* x_host, float* y_host, int32_t path,
                            int32_t* n_tensor_layers) {
    if (!c || net < 0 || net > 3 || n_cols < 1 || !x_host || !y_host) return fail(c, PDEB200_EINVAL, "net_forward: bad argument");
    const HostNet& n = c->nets[net];
    if (!n.n_layers) return fail(c, PDEB200_ESTATE, "net_forward: network not set");
    cudaSetDevice(c->device);
    const int ni = n.sizes[0], no = n.sizes[n.n_layers];
    const int ldx = pad4(ni);
    float *dx = nullptr, *dy = nullptr;
    PDEB_CUDA(c, cudaMalloc(&dx, (size_t)n_cols * ldx * sizeof(float)));
    PDEB_CUDA(c, cudaMalloc(&dy, (size_t)n_cols * no * sizeof(float)));
    PDEB_CUDA(c, cudaMemsetAsync(dx, 0, (size_t)n_cols * ldx * sizeof(float), c->stream));
    PDEB_CUDA(c, cudaMemcpy2DAsync(dx, (size_t)ldx * 4, x_host, (size_t)ni * 4, (size_t)ni * 4, n_cols, cudaMemcpyHostToDevice, c->stream));
    int32_t rc = pdeb200_net_forward_device(c, net, n_cols, dx, ldx, dy, no, path, n_tensor_layers);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(y_host, dy, (size_t)n_cols * no * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, PDEB200_ECUDA, std::string("net_forward: ") + cudaGetErrorString(e));
    }
    cudaFree(dx); cudaFree(dy);
    return rc;
}

}  // extern "C"
