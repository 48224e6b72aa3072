// Network operator entry points: the callable CustomNeuralNetworkApproximator of the reference
// (src/custom_nna.jl:13, `(app)(x) = app.model(x)` on a (rows, n_columns) matrix) for arbitrary column counts,
// layer by layer:  dense layers (in, out >= 32) on the tensor cores (dense_tc.cuh, 3xTF32 tcgen05), thin ones
// (the K <= 13 input layers and the 1-wide output heads of the shipped networks) on CUDA cores.
#include <algorithm>
#include <vector>

#include "ctx.hpp"
#include "dense_tc.cuh"

namespace pdeb200 {
namespace {

inline int pad4(int n) { return (n + 3) / 4 * 4; }

// Wt[n][k] = W[n + N*k]  (Flux Dense weight (out, in) column-major -> K-major rows, leading dimension ldw)
__global__ void transpose_w_kernel(int N, int K, int ldw, const float* __restrict__ W, float* __restrict__ Wt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * ldw) return;
    const int n = i / ldw, k = i % ldw;
    Wt[i] = k < K ? W[n + (size_t)N * k] : 0.f;
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
        int32_t rc = ensure(c, &g_scr.wt, &g_scr.wt_cap, (size_t)N * ldw);
        if (rc) return rc;
        transpose_w_kernel<<<(N * ldw + 255) / 256, 256, 0, c->stream>>>(N, K, ldw, W, g_scr.wt);
        tc::DenseArgs A;
        A.X = X; A.ldx = ldx; A.Wt = g_scr.wt; A.ldw = ldw; A.bias = bias; A.Y = Y; A.ldy = ldy;
        A.M = M; A.N = N; A.K = ldw; A.act = act;      // padded K columns are zero in Wt; X's are zeroed by the caller
        static thread_local bool configured = false;
        if (!configured) {
            PDEB_CUDA(c, cudaFuncSetAttribute(tc::dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
            configured = true;
        }
        const dim3 grid((M + tc::BM - 1) / tc::BM, (N + tc::BN - 1) / tc::BN);
        tc::dense_tc_kernel<<<grid, 160, tc::SMEM_BYTES, c->stream>>>(A);
        PDEB_CUDA(c, cudaGetLastError());
        c->launches += 2;
        if (used_tc) *used_tc += 1;
        return PDEB200_OK;
    }
    const long long total = (long long)M * N;
    dense_thin_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(M, K, N, X, ldx, W, bias, act, Y, ldy);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
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
