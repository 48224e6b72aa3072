// Dense layer on the 5th-generation tensor cores:  Y[M x N] = act(X[M x K] * Wt[N x K]^T + b)   (fp32 in / out)
//
// The weight-shared per-actuator networks of the reference (Flux Chain of Dense applied to a (ns, n_columns)
// matrix, src/PDEagent.jl:18-44, src/custom_nna.jl:13) are a 1x1 convolution over the column axis, i.e. a GEMM
// with M = n_columns (up to envs x actuators ~ 5e5).  For the shipped networks the contraction is far too thin
// (K <= 13) and the CUDA-core kernels are used; this kernel serves the dense case `drop_middle_layer = false`
// (hidden x hidden layers, 140..340 wide).
//
// Numerics: the networks are Float32 and parity is asserted at 1e-5, which a plain TF32 product (10-bit mantissa)
// misses.  Operands are therefore split on the fly, x = hi + lo with hi = x truncated to TF32, and three
// tcgen05.mma (hi*hi, hi*lo, lo*hi) accumulate into the same fp32 TMEM tile ("3xTF32": relative error ~2^-21).
//
// Structure (one 128 x 128 output tile per CTA, 9 warps; BK = 16 / SWIZZLE_64B with two CTAs per SM was measured
// and is slower: twice the barrier round trips per tile):
//   warps 0-7  producers: the weight operand (pre-split hi / lo arrays) comes by TMA (cp.async.bulk.tensor.2d through
//              SWIZZLE_128B tensor maps); the activation operand: coalesced 16-byte global loads -> hi/lo split in registers -> st.shared into the
//              K-major SWIZZLE_128B canonical layout (the split needs the data in registers, hence no TMA here);
//              register double-buffered (the next k-block's loads fly while this one is split and stored), 3-stage ring, full/empty mbarriers; afterwards the same warps run the epilogue
//              (tcgen05.ld 32 lanes x 32 columns -> transpose through the now idle stage memory -> bias + activation
//              [x act'(mask)] -> coalesced 256-byte row stores).
//   warp 8     TMEM allocation; one elected lane issues tcgen05.mma (M = 128, N = 128, K = 8 per instruction,
//              kind::tf32, both operands K-major from shared-memory descriptors) and tcgen05.commit.
#pragma once
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdeb200 {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int ROW_BYTES = BK * 4;                          // 128 B rows -> SWIZZLE_128B (BK = 16: 64 B rows -> SWIZZLE_64B)
constexpr int CHUNKS = ROW_BYTES / 16;                     // 16-byte chunks per tile row
constexpr int ATOM_BYTES = 8 * ROW_BYTES;                  // 8-row swizzle atom = stride between row groups (SBO)
constexpr int TILE_BYTES = BM * ROW_BYTES;                 // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;                // A_hi, A_lo, B_hi, B_lo
constexpr int N_PRODUCERS = 256;                           // 8 producer warps + 1 MMA warp + 4 epilogue warps
constexpr int N_EPILOGUE = 128;
constexpr int N_THREADS = N_PRODUCERS + 32 + N_EPILOGUE;
constexpr int EPI_BYTES = 4 * 32 * 33 * 4 + 128;           // epilogue transpose tiles (+ pad to keep the barriers 8-aligned)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct DenseArgs {
    // TRANSPOSED = false:  Y[M x N] = epilogue(X[M x K] * Wt[N x K]^T)
    //   X  [M][K] row-major, ldx % 4 == 0;  Wt [N][K] row-major (K-major), ldw % 4 == 0;  K % 4 == 0
    // TRANSPOSED = true :  Y_z[M x N] = sum over the z-th slice of the contraction index k of  X[k][M]^T * Wt[k][N]
    //   (both operands stored with the CONTRACTION index as the row: weight gradients dW = delta^T * input,
    //    contraction over the batch; split along it over gridDim.z, partial tiles at Y + z * y_split_stride)
    const float* X; long long ldx;
    const float* Wt; long long ldw;
    const float* bias;                 // [N] or nullptr
    float* Y; long long ldy;           // [M][N] row-major
    int M, N, K;
    int act;                           // 0 identity, 1 relu, 2 tanh  (applied to acc + bias)
    const float* mask; long long ldm;  // optional: Y *= act'(mask[m][n]) with mask_act (backward through the previous layer)
    int mask_act;
    int split_len;                     // TRANSPOSED: contraction elements per z-slice (multiple of BK)
    long long y_split_stride;
};

// B_TMA: the second operand (weights) is pre-split in global memory into hi / lo arrays and staged by TMA
// (cp.async.bulk.tensor.2d, SWIZZLE_128B tensor maps, zero fill outside [N][K]); only the activation operand is
// register staged.
struct DenseTmaMaps {
    CUtensorMap w_hi, w_lo;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}

// K-major swizzled canonical layout (cute::UMMA K-major "B64" / "B128"): rows of ROW_BYTES, 8-row groups
// ATOM_BYTES apart (SBO); the 16-byte chunk c of row r is stored at chunk
//   c ^ ((r >> 1) & 3)  for 64-byte rows  (Swizzle<2,4,3>: address bits [7,9) xor-ed into bits [4,6))
//   c ^ (r & 7)         for 128-byte rows (Swizzle<3,4,3>).
// Descriptor fields as in cute::UMMA::SmemDescriptor (sm_100).
__device__ __forceinline__ int swizzled_offset(int row, int c) {
    const int x = ROW_BYTES == 64 ? ((row >> 1) & 3) : (row & 7);
    return (row >> 3) * ATOM_BYTES + (row & 7) * ROW_BYTES + ((c ^ x) << 4);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);           // start address
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(ATOM_BYTES >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)(ROW_BYTES == 64 ? 4 : 2) << 61;  // SWIZZLE_64B / SWIZZLE_128B
    return d;
}

// instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t instr_desc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float act_f(int kind, float v) {
    if (kind == 1) return v > 0.f ? v : 0.f;
    if (kind == 2) return tanhf(v);
    return v;
}

// split a float4 into TF32-representable hi and the remainder lo
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
}

constexpr int ITS = BM * CHUNKS / N_PRODUCERS;  // 16-byte chunks per producer thread and tile
constexpr int KT = BK / 2;                      // contraction elements per thread of the transposing producer

// global -> registers (coalesced 16-byte loads); the split + st.shared happens one k-block later (store_tile), so a
// k-block's loads are in flight while the previous one is converted and the ring slot is awaited
__device__ __forceinline__ void fetch_tile(const float* __restrict__ src, long long ld, int row0, int n_rows, int k0, int K,
                                           float4 (&v)[ITS], int tid) {
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
        const int q = it * N_PRODUCERS + tid, row = q / CHUNKS, c = q % CHUNKS;
        const int gr = row0 + row, gk = k0 + c * 4;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < n_rows && gk < K) v[it] = __ldg(reinterpret_cast<const float4*>(src + (long long)gr * ld + gk));
    }
}
__device__ __forceinline__ void store_tile(const float4 (&v)[ITS], unsigned char* hi_tile, unsigned char* lo_tile, int tid) {
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
        const int q = it * N_PRODUCERS + tid, row = q / CHUNKS, c = q % CHUNKS;
        const int off = swizzled_offset(row, c);
        float4 hi, lo;
        split4(v[it], hi, lo);
        *reinterpret_cast<float4*>(hi_tile + off) = hi;
        *reinterpret_cast<float4*>(lo_tile + off) = lo;
    }
}

// Transposing producer: tile row r <-> column (col0 + r) of a row-major array whose ROWS are the contraction index.
// Thread tid owns half of tile row (tid & 127): KT coalesced scalar loads (consecutive threads read consecutive
// addresses), then KT/4 swizzled 16-byte chunks of the row.
__device__ __forceinline__ void fetch_tile_t(const float* __restrict__ src, long long ld, int col0, int n_cols, int k0, int k_end,
                                             float (&v)[KT], int tid) {
    const int row = tid & 127, half = tid >> 7;
    const int gc = col0 + row;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        const int gk = k0 + KT * half + j;
        v[j] = (gc < n_cols && gk < k_end) ? __ldg(src + (long long)gk * ld + gc) : 0.f;
    }
}
__device__ __forceinline__ void store_tile_t(const float (&v)[KT], unsigned char* hi_tile, unsigned char* lo_tile, int tid) {
    const int row = tid & 127, half = tid >> 7;
#pragma unroll
    for (int cc = 0; cc < KT / 4; ++cc) {
        const int c = (KT / 4) * half + cc;
        const int off = swizzled_offset(row, c);
        float4 hi, lo;
        split4(make_float4(v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]), hi, lo);
        *reinterpret_cast<float4*>(hi_tile + off) = hi;
        *reinterpret_cast<float4*>(lo_tile + off) = lo;
    }
}

template <bool TRANSPOSED> struct ProducerRegs;
template <> struct ProducerRegs<false> { float4 a[ITS], b[ITS]; };
template <> struct ProducerRegs<true> { float a[KT], b[KT]; };

template <bool TRANSPOSED, bool B_TMA>
__device__ __forceinline__ void producer_fetch(const DenseArgs& A, int m0, int n0, int k0, int k_end, ProducerRegs<TRANSPOSED>& R, int tid) {
    if constexpr (TRANSPOSED) {
        fetch_tile_t(A.X, A.ldx, m0, A.M, k0, k_end, R.a, tid);
        fetch_tile_t(A.Wt, A.ldw, n0, A.N, k0, k_end, R.b, tid);
    } else {
        fetch_tile(A.X, A.ldx, m0, A.M, k0, k_end, R.a, tid);
        if constexpr (!B_TMA) fetch_tile(A.Wt, A.ldw, n0, A.N, k0, k_end, R.b, tid);
    }
}
template <bool TRANSPOSED, bool B_TMA>
__device__ __forceinline__ void producer_store(const ProducerRegs<TRANSPOSED>& R, unsigned char* st, int tid) {
    if constexpr (TRANSPOSED) {
        store_tile_t(R.a, st, st + TILE_BYTES, tid);
        store_tile_t(R.b, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, tid);
    } else {
        store_tile(R.a, st, st + TILE_BYTES, tid);
        if constexpr (!B_TMA) store_tile(R.b, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, tid);
    }
}

__device__ __forceinline__ float act_grad_f(int kind, float out) {
    if (kind == 1) return out > 0.f ? 1.f : 0.f;
    if (kind == 2) return 1.f - out * out;
    return 1.f;
}

__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// Persistent kernel: one CTA per SM walks the output tiles (tile = blockIdx.x, += gridDim.x; the N tiles of an M tile
// are neighbours so that the activation tile is shared through L2).  Roles:
//   warps 0-7   producers (activation operand in registers -> hi/lo split -> swizzled stage; thread 0 also issues the
//               weight TMA)
//   warp  8     MMA issuer; the fp32 accumulator is DOUBLE BUFFERED in TMEM (2 x 128 columns): the MMAs of tile i + 1
//               run while the epilogue warps drain tile i
//   warps 9-12  epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = 32 output rows; per 32-column chunk
//               tcgen05.ld -> transpose through a private 32 x 33 shared-memory tile -> bias / activation / act'(mask)
//               -> 128-byte coalesced row stores
template <bool TRANSPOSED, bool B_TMA>
__global__ void __launch_bounds__(N_THREADS, 1) dense_tc_kernel(const __grid_constant__ DenseArgs A,
                                                                const __grid_constant__ DenseTmaMaps TM) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = s32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                   // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* tiles = smem_raw + (base - raw);
    float* epi_buf = reinterpret_cast<float*>(tiles + STAGES * STAGE_BYTES);                 // [4 warps][32][33]
    const uint32_t bars = base + STAGES * STAGE_BYTES + EPI_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tiles + STAGES * STAGE_BYTES + EPI_BYTES + 16 * STAGES + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (A.M + BM - 1) / BM, tiles_n = (A.N + BN - 1) / BN;
    const int n_z = TRANSPOSED ? (A.K + A.split_len - 1) / A.split_len : 1;
    const int n_tiles = tiles_m * tiles_n * n_z;
    constexpr int MMA_WARP = N_PRODUCERS / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { bar_init(full0 + 8 * s, N_PRODUCERS + (B_TMA ? 1 : 0)); bar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { bar_init(tfull0 + 8 * b, 1); bar_init(tempty0 + 8 * b, N_EPILOGUE); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // tile -> (m tile, n tile, contraction slice)
    auto decode = [&](int tile, int& m0, int& n0, int& z, int& k_begin, int& k_end) {
        const int tn = tile % tiles_n, rest = tile / tiles_n;
        const int tm = rest % tiles_m;
        z = rest / tiles_m;
        m0 = tm * BM; n0 = tn * BN;
        k_begin = TRANSPOSED ? z * A.split_len : 0;
        k_end = TRANSPOSED ? min(A.K, k_begin + A.split_len) : A.K;
    };

    if (warp < MMA_WARP) {
        // ===== producers =====
        const int tid = threadIdx.x;
        ProducerRegs<TRANSPOSED> R0, R1;
        int it = 0;                                                     // k-blocks issued so far (ring position)
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int m0, n0, z, k_begin, k_end;
            decode(tile, m0, n0, z, k_begin, k_end);
            const int KB = (k_end - k_begin + BK - 1) / BK;
            auto commit = [&](int kb, const ProducerRegs<TRANSPOSED>& R) {
                const int g = it + kb, s = g % STAGES;
                bar_wait(empty0 + 8 * s, ((g / STAGES) & 1) ^ 1);
                if constexpr (B_TMA) {
                    if (tid == 0) {     // weights: two 128-row x 128-byte boxes (hi, lo) straight into the swizzled stage
                        bar_expect_tx(full0 + 8 * s, 2 * TILE_BYTES);
                        tma_load_2d(base + s * STAGE_BYTES + 2 * TILE_BYTES, &TM.w_hi, k_begin + kb * BK, n0, full0 + 8 * s);
                        tma_load_2d(base + s * STAGE_BYTES + 3 * TILE_BYTES, &TM.w_lo, k_begin + kb * BK, n0, full0 + 8 * s);
                    }
                }
                producer_store<TRANSPOSED, B_TMA>(R, tiles + s * STAGE_BYTES, tid);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor-core reads
                bar_arrive(full0 + 8 * s);
            };
            // software pipeline over two register sets: the loads of k-block kb + 1 fly while kb is split and stored
            if (KB > 0) producer_fetch<TRANSPOSED, B_TMA>(A, m0, n0, k_begin, k_end, R0, tid);
            for (int kb = 0; kb < KB; kb += 2) {
                if (kb + 1 < KB) producer_fetch<TRANSPOSED, B_TMA>(A, m0, n0, k_begin + (kb + 1) * BK, k_end, R1, tid);
                commit(kb, R0);
                if (kb + 1 < KB) {
                    if (kb + 2 < KB) producer_fetch<TRANSPOSED, B_TMA>(A, m0, n0, k_begin + (kb + 2) * BK, k_end, R0, tid);
                    commit(kb + 1, R1);
                }
            }
            it += KB;
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        const uint32_t idesc = instr_desc();
        int it = 0, lt = 0;                                             // ring position, local tile counter
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            int m0, n0, z, k_begin, k_end;
            decode(tile, m0, n0, z, k_begin, k_end);
            const int KB = (k_end - k_begin + BK - 1) / BK;
            const int buf = lt & 1;
            bar_wait(tempty0 + 8 * buf, ((lt >> 1) & 1) ^ 1);           // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)(buf * BN);
            for (int kb = 0; kb < KB; ++kb) {
                const int g = it + kb, s = g % STAGES;
                bar_wait(full0 + 8 * s, (g / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        // advancing 8 tf32 (32 bytes) inside a swizzled row = +2 in the encoded start address
                        const uint64_t ahi = smem_desc(st + k * 32), alo = smem_desc(st + TILE_BYTES + k * 32);
                        const uint64_t bhi = smem_desc(st + 2 * TILE_BYTES + k * 32), blo = smem_desc(st + 3 * TILE_BYTES + k * 32);
                        mma_tf32(acc, alo, bhi, idesc, (kb | k) != 0);      // small terms first
                        mma_tf32(acc, ahi, blo, idesc, 1);
                        mma_tf32(acc, ahi, bhi, idesc, 1);
                    }
                    // commit: arrives on the barrier when the MMAs above have finished reading the stage
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * s) : "memory");
                }
                __syncwarp();
            }
            if (lane == 0)      // accumulator complete (also for an empty contraction slice: nothing pending)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tfull0 + 8 * buf) : "memory");
            __syncwarp();
            it += KB;
        }
    } else {
        // ===== epilogue warps =====
        const int lq = warp & 3;                                        // a warp may only touch TMEM lanes [32 (warp % 4), +32)
        float* tbuf = epi_buf + lq * 32 * 33;
        int lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            int m0, n0, z, k_begin, k_end;
            decode(tile, m0, n0, z, k_begin, k_end);
            const int KB = (k_end - k_begin + BK - 1) / BK;
            const int buf = lt & 1;
            float* const Yz = A.Y + (TRANSPOSED ? (long long)z * A.y_split_stride : 0);
            bar_wait(tfull0 + 8 * buf, (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + 32 >= BN) {                                     // last chunk read: hand the accumulator back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    bar_arrive(tempty0 + 8 * buf);
                }
                if (n0 + c0 >= A.N) continue;                            // whole chunk outside N (warp-uniform)
                const int n = n0 + c0 + lane;
                const bool nv = n < A.N;
                const float bv = (A.bias && nv) ? __ldg(A.bias + n) : 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = KB > 0 ? __uint_as_float(r[j]) : 0.f;
                __syncwarp();
                if (A.mask) {
                    // backward epilogue, 16 rows at a time: the 16 act'(.) inputs are LOADED first (unconditional,
                    // clamped addresses) and only then consumed -- the pipeline issues in order, so a load followed
                    // directly by its use would expose one full memory latency per row
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        float mk[16];
                        const int nc = nv ? n : 0;
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const int row = min(m0 + lq * 32 + h * 16 + rr, A.M - 1);
                            mk[rr] = __ldg(A.mask + (long long)row * A.ldm + nc);
                        }
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const int row = m0 + lq * 32 + h * 16 + rr;
                            if (row < A.M && nv)
                                Yz[(long long)row * A.ldy + n] =
                                    act_f(A.act, tbuf[(h * 16 + rr) * 33 + lane] + bv) * act_grad_f(A.mask_act, mk[rr]);
                        }
                    }
                } else {
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = m0 + lq * 32 + rr;
                        if (row >= A.M) break;
                        if (nv) Yz[(long long)row * A.ldy + n] = act_f(A.act, tbuf[rr * 33 + lane] + bv);
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN) : "memory");
    }
}

}  // namespace tc
}  // namespace pdeb200
