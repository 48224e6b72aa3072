// One-shot gradient exchange over NVLink peer memory (the path's only collective).
//
// The reference is single-process (no collective exists in it); under data parallelism over environments the
// DDPG update (src/PDEagent.jl:363-418) needs the sum of every rank's [critic | actor] gradient.  That message is
// 2-22 KB: latency-bound, so it does not go through a ring/tree collective launched from the host but is executed
// by the LAST CTA of the gradient kernel itself: it stores its reduced gradient into a per-source slot of every
// peer's exchange buffer (plain stores through NVLink / NVSwitch to cudaIpc-mapped peer memory), publishes an epoch
// flag with release.sys, waits for the peers' flags with acquire.sys, and sums the slots in ascending rank order --
// every rank performs the identical fixed-order sum, so weights stay bit-identical across ranks.  ADAM and the
// Polyak step follow in the same CTA.  No extra launch, no host involvement, CUDA-graph capturable.
//
// Buffer (one per rank, cudaMalloc + cudaIpcGetMemHandle, opened by every peer):
//   [0, 64)        uint32 gflag[2][8]    epoch stamps of the gradient exchange, [parity][source rank]
//   [64, 128)      uint32 sflag[2][8]    same for the batch-statistics exchange
//   [128, 1152)    double sdata[2][8][8] batch statistics {sum r, sum r^2, n}
//   [4096, ...)    uint64 ldata[2][8][cap/2]   gradient exchange, one {epoch : float} word per element (flag-in-data, below)
// Sliced exchange (the one-cluster update kernel, ddpg_fast.cuh): the 16 CTAs of the cluster each own 1/16 of the parameter
// vector; CTA s exchanges its slice at offset s * cap/32 inside the same per-source slot, all 16 slices in flight at once.
// Two parities suffice: a rank that writes exchange e+2 has passed the wait of e+1, i.e. has seen every peer's
// e+1 flag, which each peer stores only after it has finished reading exchange e.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdeb200 {

constexpr int kMaxRanks = 8;
constexpr size_t kCommHeaderBytes = 4096;
constexpr int kMaxSlices = 16;

struct CommDev {
    int rank = 0, nranks = 1;           // nranks <= 1: no exchange
    int cap = 0;                        // floats per gradient slot
    unsigned int* epoch = nullptr;      // device counters (own memory): [0] gradient exchanges, [1] statistics exchanges
    int* err = nullptr;                 // set to 1 when a wait timed out (a peer died): results are invalid, no hang
    unsigned long long timeout_ns = 0;
    char* peer[kMaxRanks] = {nullptr};  // exchange buffer of every rank as mapped into THIS process (own = local)
};

__device__ __forceinline__ unsigned int* comm_gflag(char* base, int par, int src) { return (unsigned int*)base + par * kMaxRanks + src; }
__device__ __forceinline__ unsigned int* comm_sflag(char* base, int par, int src) { return (unsigned int*)(base + 64) + par * kMaxRanks + src; }
__device__ __forceinline__ double* comm_sdata(char* base, int par, int src) { return (double*)(base + 128) + (par * kMaxRanks + src) * 8; }
__device__ __forceinline__ float* comm_gdata(char* base, int cap, int par, int src) {
    return (float*)(base + kCommHeaderBytes) + (size_t)(par * kMaxRanks + src) * cap;
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void comm_wait_flag(const CommDev& cm, const unsigned int* flag, unsigned int e) {
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) != e) {
        if (global_timer_ns() - t0 > cm.timeout_ns) { *cm.err = 1; break; }
        __nanosleep(64);
    }
}

// ---- flag-in-data ("LL") transport of the gradient exchanges ------------------------------------------------------------------
// A separate flag needs  data stores -> system-scope release (a round trip to the peer: every store acknowledged) -> flag
// store -> peer polls: measured 5 us per exchange on 2 B200s.  Here every element travels as ONE 64-bit store
// {epoch : float bits} (single-copy atomic), so the receiver polls the data words themselves: one one-way NVLink latency.
// The gdata region is reinterpreted as uint64 [2][8][cap/2].
__device__ __forceinline__ unsigned long long* comm_ldata(char* base, int cap, int par, int src) {
    return (unsigned long long*)(base + kCommHeaderBytes) + (size_t)(par * kMaxRanks + src) * (cap / 2);
}
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned int e) {
    const unsigned long long w = ((unsigned long long)e << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ float ll_wait(const CommDev& cm, const unsigned long long* p, unsigned int e) {
    unsigned long long w;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
        if ((unsigned int)(w >> 32) == e) break;
        if (global_timer_ns() - t0 > cm.timeout_ns) { *cm.err = 1; break; }
    }
    return __uint_as_float((unsigned int)w);
}

// Sum of vec[0..n) over all ranks by the threads of ONE CTA per rank; out may alias vec.  n <= cap/2 (whole-vector form,
// offset 0) or cap/2/kMaxSlices (sliced form: 16 CTAs of the one-cluster kernel, each with its own offset).  Fixed rank-order
// sum in double => identical results on every rank.
__device__ __forceinline__ void comm_allreduce_ll(const CommDev& cm, unsigned int e, int off, const float* vec, float* out, int n) {
    const int par = (int)(e & 1u);
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const float v = vec[q];
#pragma unroll 1
        for (int p = 0; p < cm.nranks; ++p) ll_store(comm_ldata(cm.peer[p], cm.cap, par, cm.rank) + off + q, v, e);
    }
    char* own = cm.peer[cm.rank];
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < cm.nranks; ++r) s += (double)ll_wait(cm, comm_ldata(own, cm.cap, par, r) + off + q, e);
        out[q] = (float)s;
    }
    __syncthreads();
}

// whole-vector form for the last-CTA tails (two-launch update kernels): epoch from the device counter
__device__ __forceinline__ void comm_allreduce_cta(const CommDev& cm, const float* vec, float* out, int n) {
    __shared__ unsigned int s_epoch;
    if (threadIdx.x == 0) s_epoch = cm.epoch[0] + 1;
    __syncthreads();
    const unsigned int e = s_epoch;
    comm_allreduce_ll(cm, e, 0, vec, out, n);
    if (threadIdx.x == 0) cm.epoch[0] = e;
    __syncthreads();
}

// sliced form: e = epoch stamp (read from cm.epoch[0] at kernel start + phase; the kernel's rank-0 CTA stores the final value
// back); slice s uses offset s * cap / 2 / kMaxSlices
__device__ __forceinline__ void comm_allreduce_slice(const CommDev& cm, unsigned int e, int slice, const float* vec, float* out, int n) {
    comm_allreduce_ll(cm, e, slice * (cm.cap / 2 / kMaxSlices), vec, out, n);
}

// The batch statistics {sum r, sum r^2, n} summed over all ranks, executed by ONE thread per rank.
__device__ __forceinline__ void comm_allreduce_stats(const CommDev& cm, double* v3) {
    const unsigned int e = cm.epoch[1] + 1;
    const int par = (int)(e & 1u);
    for (int p = 0; p < cm.nranks; ++p) {
        double* d = comm_sdata(cm.peer[p], par, cm.rank);
        d[0] = v3[0]; d[1] = v3[1]; d[2] = v3[2];
    }
    __threadfence_system();
    for (int p = 0; p < cm.nranks; ++p) st_release_sys(comm_sflag(cm.peer[p], par, cm.rank), e);
    char* own = cm.peer[cm.rank];
    double a = 0.0, b = 0.0, n = 0.0;
    for (int r = 0; r < cm.nranks; ++r) {
        comm_wait_flag(cm, comm_sflag(own, par, r), e);
        const double* d = comm_sdata(own, par, r);
        a += __ldcg(d); b += __ldcg(d + 1); n += __ldcg(d + 2);
    }
    v3[0] = a; v3[1] = b; v3[2] = n;
    cm.epoch[1] = e;
}

}  // namespace pdeb200
