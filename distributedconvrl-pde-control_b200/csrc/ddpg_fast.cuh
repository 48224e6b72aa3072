// Register-resident DDPG update for the shipped network shapes -- included by agent.cu only (inside its unnamed
// namespace, after Ring / AgentDev / philox_u64 / ordered_sum_cg / CommDev).
//
// Restates update!(policy, batch) (/root/reference/src/PDEagent.jl:363-418) and, in fetch mode, pde_sample / pde_fetch!
// (:317-340) for two-layer networks  actor ns -> ha -> 1,  critic (ns+1) -> hc -> 1  (create_NNA with
// drop_middle_layer = true, PDEagent.jl:14-56 -- every shipped script: KS 1-6-1 / 2-140-1, Keller-Segel 12-20-1 / 13-340-1,
// Fluid 9-18-1 / 10-340-1).
//
// Why a second implementation: the runtime-shaped shared-memory kernels above run ~18 barrier-separated phases per
// 32-sample tile at ~10 % issue utilisation (profiles/r1_ddpg.md: 70 us per update in 3 launches).  Here
//   * a WARP owns a tile of 32 samples; its LANES own hidden units (UPL per lane, WC warps per tile when the hidden
//     layer is wider than 32*UPL), with that unit's weights AND gradient accumulators in registers for the whole kernel:
//     the gradient never needs a cross-lane reduction;
//   * the only cross-lane traffic is the per-sample output  q_i = sum_units w2_u h_u(i): the per-lane partials of 8
//     samples at a time are reduced with a 7-shuffle transpose-reduce + 2 butterfly rounds that leave q_i on lane i,
//     where the sample's scalars (r, t, target) live;
//   * quirk Q1's batch-mean reward r-bar would make the critic gradient wait for a grid-wide (and cross-rank) sum of the
//     sampled rewards; the gradient is LINEAR in r-bar, so the kernel accumulates G_A = sum_i (T_i - q_i) dq_i/dtheta
//     and G_B = sum_i dq_i/dtheta and the tail forms  g = -(2/n) (G_A + r-bar G_B)  after ONE exchange that carries
//     G_A, G_B and {sum r, sum r^2, n, sum c, sum c^2}: the sampler's separate statistics exchange disappears;
//   * the sampler itself (Philox draw + ring gather) is fused into the critic kernel (fetch mode): one update = TWO
//     launches, each ending in the last-CTA tail (fixed-order reduction, NVLink peer exchange, ADAM, Polyak).
#pragma once

constexpr int kXTail = 8;            // scalars appended to the exchange vector: sum r, sum r^2, n, sum c, sum c^2 / sum q, n

struct FastArgs {
    const float *pA, *pC, *pAt, *pCt;        // flat parameter vectors (Flux layout, see NetDev)
    int ns, ha, hc, actA2;                   // state rows, hidden widths, actor output activation (hidden: relu)
    int nA, nC;                              // parameter counts
    int batch;
    float *bs, *ba, *br, *bs2; uint8_t* bt; int64_t* inds;     // staged batch: written in fetch mode, read otherwise
    int fetch; unsigned long long seed; AgentDev* dev; long long ncols;
    const float *rstate, *raction, *rreward; const uint8_t* rterminal;
    float gamma; int literal;
    float* partials; int n_x;                // [gridDim.x][n_x]
    unsigned int* ticket; float* xbuf; float* grads; double* stats; float* losses;
    float *x, *m, *v, *target; double* betap; double eta, b1, b2, eps; float polyak;
    CommDev cm;
    unsigned long long* tl;                  // optional timeline (PDEB200_DDPG_TIMELINE=1): [0] = record count, then 6 words per kernel
};

// timeline record of one kernel: {phase, t(entry of CTA 0), t(CTA 0 loop done), t(tail start), t(exchange done), t(tail end)} in ns
__device__ __forceinline__ void tl_mark(const FastArgs& F, int slot) {
    if (F.tl && threadIdx.x == 0) F.tl[4096 + slot] = global_timer_ns();
}

// Samples are processed in groups of GS = 8 inside NON-unrolled loops: the first version unrolled all 32 samples of every
// pass (12.7k SASS lines executed exactly once per warp) and ncu attributed 49 % of the critic kernel's samples to
// instruction-cache misses (stall_no_inst, profiles/r2_ddpg.md).
constexpr int GS = 8;

// v[k] = this lane's partial for sample k of the group.  Returns, on EVERY lane, the sum over the 32 lanes for sample
// (lane & 7): a 3-round transpose-reduce (7 shuffles) followed by two butterfly rounds.
__device__ __forceinline__ float group_reduce(float (&v)[GS], int lane) {
#pragma unroll
    for (int s = GS / 2; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const float send = up ? v[k] : v[k + s];
            const float keep = up ? v[k + s] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    float r = v[0];
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// A two-layer network NIN -> nh (relu) -> 1 with the hidden units spread over lanes: lane owns units ubase + 32 k + lane.
template <int NIN, int UPL>
struct UnitNet {
    float w1[UPL][NIN], b1[UPL], w2[UPL], b2;

    // L2 = true: read through L2 (ld.global.cg) -- the one-cluster kernel re-reads critic weights that OTHER CTAs of the
    // cluster have just updated; a non-coherent / L1-cached load could return the line this SM fetched before the update
    template <bool L2 = false>
    __device__ __forceinline__ void load(const float* p, int nh, int ubase, int lane) {
        auto ld = [](const float* q) { return L2 ? __ldcg(q) : __ldg(q); };
#pragma unroll
        for (int k = 0; k < UPL; ++k) {
            const int u = ubase + 32 * k + lane;
            const bool ok = u < nh;
#pragma unroll
            for (int j = 0; j < NIN; ++j) w1[k][j] = ok ? ld(p + u + nh * j) : 0.f;
            b1[k] = ok ? ld(p + nh * NIN + u) : 0.f;
            w2[k] = ok ? ld(p + nh * NIN + nh + u) : 0.f;             // 0 for padding units: no forward or backward contribution
        }
        b2 = ld(p + nh * NIN + 2 * nh);
    }
    __device__ __forceinline__ float hidden(int k, const float (&x)[NIN]) const {   // relu.(W x .+ b), bias added last
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < NIN; ++j) acc = fmaf(w1[k][j], x[j], acc);
        return fmaxf(acc + b1[k], 0.f);
    }
    // returns on lane i the sum over this warp's units of w2_u h_u(i); rows of the tile in shared memory xs[i * XS + j]
    template <int XS, bool STORE>
    __device__ __forceinline__ float forward(const float* xs, float* hs, int hp, int ubase, int lane) const {
        float out = 0.f;
#pragma unroll 1
        for (int g = 0; g < 32 / GS; ++g) {
            float v[GS];
#pragma unroll
            for (int ii = 0; ii < GS; ++ii) {
                const int i = g * GS + ii;
                float x[NIN];
#pragma unroll
                for (int j = 0; j < NIN; ++j) x[j] = xs[i * XS + j];
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < UPL; ++k) {
                    const float h = hidden(k, x);
                    if (STORE) hs[i * hp + ubase + 32 * k + lane] = h;
                    acc = fmaf(w2[k], h, acc);
                }
                v[ii] = acc;
            }
            const float r = group_reduce(v, lane);
            if ((lane / GS) == g) out = r;
        }
        return out;
    }
};

template <int NIN, int UPL, int NG>
struct UnitGrads {
    float w1[NG][UPL][NIN], b1[NG][UPL], w2[NG][UPL];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int k = 0; k < UPL; ++k) {
                b1[g][k] = 0.f; w2[g][k] = 0.f;
#pragma unroll
                for (int j = 0; j < NIN; ++j) w1[g][k][j] = 0.f;
            }
    }
    // own units' entries of the flat gradient (Flux layout) into dst + g * stride_g
    __device__ __forceinline__ void store(float* dst, int stride_g, int nh, int ubase, int lane) const {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int k = 0; k < UPL; ++k) {
                const int u = ubase + 32 * k + lane;
                if (u < nh) {
                    float* d = dst + g * stride_g;
#pragma unroll
                    for (int j = 0; j < NIN; ++j) d[u + nh * j] = w1[g][k][j];
                    d[nh * NIN + u] = b1[g][k];
                    d[nh * NIN + nh + u] = w2[g][k];
                }
            }
    }
};

// Backward through the two-layer net for one tile: upstream deltas d0 (and d1 for the second gradient set) per sample
// in shared memory; hidden activations recomputed (RECOMP) or read back from hs (each lane reads what it stored).
// WANT_IN: vin[i] = sum over own units of d0_i relu'(h) w2_u w1[u][JIN]  (gradient w.r.t. input column JIN).
template <int NIN, int UPL, int NG, int XS, bool RECOMP, bool WANT_IN, int JIN>
__device__ __forceinline__ float unit_backward(const UnitNet<NIN, UPL>& N, UnitGrads<NIN, UPL, NG>& G, const float* xs, const float* hs, int hp,
                                               int ubase, int lane, const float* d0, const float* d1) {
    float out = 0.f;
#pragma unroll 1
    for (int g = 0; g < 32 / GS; ++g) {
        float vin[GS];
#pragma unroll
        for (int ii = 0; ii < GS; ++ii) {
            const int i = g * GS + ii;
            float x[NIN];
#pragma unroll
            for (int j = 0; j < NIN; ++j) x[j] = xs[i * XS + j];
            const float da = d0[i];
            const float db = NG > 1 ? d1[i] : 0.f;
            float acc_in = 0.f;
#pragma unroll
            for (int k = 0; k < UPL; ++k) {
                const float h = RECOMP ? N.hidden(k, x) : hs[i * hp + ubase + 32 * k + lane];
                const float gp = h > 0.f ? N.w2[k] : 0.f;                    // relu'(h) w2_u
                const float ea = da * gp;
                G.w2[0][k] = fmaf(da, h, G.w2[0][k]);
                G.b1[0][k] += ea;
#pragma unroll
                for (int j = 0; j < NIN; ++j) G.w1[0][k][j] = fmaf(ea, x[j], G.w1[0][k][j]);
                if (NG > 1) {
                    const float eb = db * gp;
                    G.w2[NG - 1][k] = fmaf(db, h, G.w2[NG - 1][k]);
                    G.b1[NG - 1][k] += eb;
#pragma unroll
                    for (int j = 0; j < NIN; ++j) G.w1[NG - 1][k][j] = fmaf(eb, x[j], G.w1[NG - 1][k][j]);
                }
                if (WANT_IN) acc_in = fmaf(ea, N.w1[k][JIN], acc_in);
            }
            vin[ii] = acc_in;
        }
        if (WANT_IN) {
            const float r = group_reduce(vin, lane);
            if ((lane / GS) == g) out = r;
        }
    }
    return out;
}

// team reduction of a per-sample value (lane i <-> sample i) over the WC warps that share a tile's hidden units
template <int WC>
__device__ __forceinline__ float team_sum(float v, float* tq, int tw, int lane) {
    if (WC == 1) return v;
    tq[tw * 32 + lane] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WC; ++w) s += tq[w * 32 + lane];
    return s;
}

__device__ __forceinline__ float out_act(int kind, float v) { return kind == 2 ? tanhf(v) : (kind == 1 ? fmaxf(v, 0.f) : v); }
__device__ __forceinline__ float out_act_grad(int kind, float out) { return kind == 2 ? 1.f - out * out : (kind == 1 ? (out > 0.f ? 1.f : 0.f) : 1.f); }

__device__ void fast_tail(const FastArgs& F, int phase, unsigned long long rng_offset);

// shared memory carve-up (floats): per warp {xsA [32][NS+1], xsB [32][NS+1], d [2][32]}; per team {tq [2][WC][32], hs [32][HP]};
// then acc [teams][n_x]; the one-cluster kernel appends accsum [n_x] + xv [slice payload]
template <int NS, int UPLC, int WC>
struct FastGeom {
    static constexpr int XS = NS + 1;
    static constexpr int HP = WC * UPLC * 32;
    static constexpr bool RECOMP = (NS + 1) <= 4;
    static constexpr int WARP_F = 2 * 32 * XS + 64;
    static constexpr int TEAM_F = (WC > 1 ? 2 * WC * 32 : 0) + (RECOMP ? 0 : 32 * HP);
    static __host__ __device__ size_t base_floats(int n_warps, int n_x) {
        const int teams = n_warps / WC;
        return (size_t)n_warps * WARP_F + (size_t)teams * TEAM_F + (size_t)teams * n_x;
    }
    static size_t smem_bytes(int n_warps, int n_x) { return base_floats(n_warps, n_x) * sizeof(float); }
};

// ---- critic phase tiles: [sample] -> targets from (A_t, C_t) -> critic gradient sets G_A, G_B ------------------------------
// Leaves this CTA's partial (sum over its teams, fixed order) in out[0 .. F.n_x): [G_A (nC) | G_B (nC) | sum r, sum r^2, n,
// sum c, sum c^2, 0, 0, 0].  Returns the sampler's Philox offset it used (fetch mode).
template <int NS, int UPLC, int WC>
__device__ __forceinline__ unsigned long long critic_tiles(const FastArgs& F, float* sm, float* out) {
    using Geo = FastGeom<NS, UPLC, WC>;
    constexpr int XS = Geo::XS, NIN = NS + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int teams = n_warps / WC, team = warp / WC, tw = warp % WC;
    float* xs2 = sm + warp * Geo::WARP_F;                 // [s'; a']
    float* xsc = xs2 + 32 * XS;                           // [s; a]
    float* dsm = xsc + 32 * XS;                           // deltas [2][32]
    float* team_base = sm + n_warps * Geo::WARP_F + team * Geo::TEAM_F;
    float* tq0 = team_base;                               // [WC][32] x 2
    float* tq1 = team_base + (WC > 1 ? WC * 32 : 0);
    float* hs = team_base + (WC > 1 ? 2 * WC * 32 : 0);
    float* acc = sm + n_warps * Geo::WARP_F + teams * Geo::TEAM_F + team * F.n_x;

    UnitNet<NS, 1> At;  At.load(F.pAt, F.ha, 0, lane);
    UnitNet<NIN, UPLC> C;  C.load(F.pC, F.hc, tw * UPLC * 32, lane);
    UnitGrads<NIN, UPLC, 2> G; G.zero();
    float s_r = 0.f, s_r2 = 0.f, s_c = 0.f, s_c2 = 0.f, s_n = 0.f, s_da = 0.f;

    Ring sa = {0, 0, 0}, rt = {0, 0, 0};
    unsigned long long offset = 0;
    if (F.fetch) { sa = F.dev->sa; rt = F.dev->rt; offset = F.dev->rng_offset; }
    const long long range = rt.len - F.ncols;
    const int n_tiles = (F.batch + 31) / 32;
    const int stride = gridDim.x * teams;
    const int iters = (n_tiles + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
        const int tile = blockIdx.x * teams + team + it * stride;
        if (WC == 1 && tile >= n_tiles) break;            // no CTA barrier in the loop: a finished warp may leave
        const int idx = tile * 32 + lane;
        const bool valid = tile < n_tiles && idx < F.batch;
        // ---- this lane's sample -----------------------------------------------------------------------------------
        float s[NS], s2[NS], a = 0.f, r = 0.f, tf = 0.f;
#pragma unroll
        for (int j = 0; j < NS; ++j) { s[j] = 0.f; s2[j] = 0.f; }
        if (valid) {
            if (F.fetch) {
                const unsigned long long u = philox_u64(F.seed, offset + (unsigned long long)idx);
                const long long ind = (long long)__umul64hi(u, (unsigned long long)range);
                const long long ps = (sa.start + ind) % sa.cap, ps2 = (sa.start + ind + F.ncols) % sa.cap, pr = (rt.start + ind) % rt.cap;
#pragma unroll
                for (int j = 0; j < NS; ++j) { s[j] = F.rstate[ps * NS + j]; s2[j] = F.rstate[ps2 * NS + j]; }
                a = F.raction[ps]; r = F.rreward[pr];
                const uint8_t tb = F.rterminal[pr];
                tf = (float)tb;
                if (tw == 0) {
                    F.inds[idx] = ind; F.ba[idx] = a; F.br[idx] = r; F.bt[idx] = tb;
#pragma unroll
                    for (int j = 0; j < NS; ++j) { F.bs[(size_t)idx * NS + j] = s[j]; F.bs2[(size_t)idx * NS + j] = s2[j]; }
                }
            } else {
#pragma unroll
                for (int j = 0; j < NS; ++j) { s[j] = F.bs[(size_t)idx * NS + j]; s2[j] = F.bs2[(size_t)idx * NS + j]; }
                a = F.ba[idx]; r = F.br[idx]; tf = (float)F.bt[idx];
            }
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) { xs2[lane * XS + j] = s2[j]; xsc[lane * XS + j] = s[j]; }
        xsc[lane * XS + NS] = a;
        __syncwarp();
        // ---- a' = A_t(s') (every warp of the team, redundantly: the actor is a handful of FMAs) -------------------
        const float a2 = out_act(F.actA2, At.template forward<XS, false>(xs2, nullptr, 0, 0, lane) + At.b2);
        xs2[lane * XS + NS] = a2;
        __syncwarp();
        // ---- q_t = C_t([s'; a']) (target weights re-read per tile, L1 hits: keeps them out of the loop's live registers) --
        float qt;
        {
            UnitNet<NIN, UPLC> Ct;
            Ct.load(F.pCt, F.hc, tw * UPLC * 32, lane);
            qt = team_sum<WC>(Ct.template forward<XS, false>(xs2, nullptr, 0, tw * UPLC * 32, lane), tq0, tw, lane) + Ct.b2;
        }
        const float T = F.gamma * (1.f - tf) * qt;
        // ---- q = C([s; a]) -------------------------------------------------------------------------------------------
        const float q = team_sum<WC>(C.template forward<XS, !Geo::RECOMP>(xsc, hs, Geo::HP, tw * UPLC * 32, lane), tq1, tw, lane) + C.b2;
        // literal (quirk Q1): loss = mean_{i,j} (r_j + T_i - q_i)^2  =>  dq_i = -(2/n)(rbar + c_i), c_i = T_i - q_i
        // per-sample:         loss = mean_i (r_i + T_i - q_i)^2      =>  dq_i = -(2/n) c_i,        c_i = r_i + T_i - q_i
        const float c = valid ? (F.literal ? T - q : r + T - q) : 0.f;
        dsm[lane] = c; dsm[32 + lane] = valid ? 1.f : 0.f;
        if (valid) { s_r += r; s_r2 = fmaf(r, r, s_r2); s_c += c; s_c2 = fmaf(c, c, s_c2); s_n += 1.f; s_da += c; }
        __syncwarp();
        unit_backward<NIN, UPLC, 2, XS, Geo::RECOMP, false, 0>(C, G, xsc, hs, Geo::HP, tw * UPLC * 32, lane, dsm, dsm + 32);
        __syncwarp();
    }
    // ---- per-team accumulators -> CTA partial (fixed order over the teams) --------------------------------------------
    if (blockIdx.x == 0) tl_mark(F, 1);
    G.store(acc, F.nC, F.hc, tw * UPLC * 32, lane);
    s_r = warp_sum(s_r); s_r2 = warp_sum(s_r2); s_c = warp_sum(s_c); s_c2 = warp_sum(s_c2); s_n = warp_sum(s_n); s_da = warp_sum(s_da);
    if (tw == 0 && lane == 0) {
        acc[F.nC - 1] = s_da;                               // b2 gradient, set A: sum_i c_i
        acc[2 * F.nC - 1] = s_n;                            //              set B: sum_i 1
        float* t = acc + 2 * F.nC;
        t[0] = s_r; t[1] = s_r2; t[2] = s_n; t[3] = s_c; t[4] = s_c2; t[5] = 0.f; t[6] = 0.f; t[7] = 0.f;
    }
    __syncthreads();
    const float* acc0 = sm + n_warps * Geo::WARP_F + teams * Geo::TEAM_F;
    for (int q = threadIdx.x; q < F.n_x; q += blockDim.x) {
        float sum = 0.f;
        for (int t = 0; t < teams; ++t) sum += acc0[t * F.n_x + q];
        out[q] = sum;
    }
    return offset;
}

// ---- actor phase tiles: gradient of -mean C([s; A(s)]) w.r.t. the actor parameters (PDEagent.jl:402-409) ------------------------
// out[0 .. F.n_x): [G (nA) | sum q, n, 0 ...].  L2: read the critic weights through L2 (see UnitNet::load).
template <int NS, int UPLC, int WC, bool L2>
__device__ __forceinline__ void actor_tiles(const FastArgs& F, float* sm, float* out) {
    using Geo = FastGeom<NS, UPLC, WC>;
    constexpr int XS = Geo::XS, NIN = NS + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int teams = n_warps / WC, team = warp / WC, tw = warp % WC;
    float* xsa = sm + warp * Geo::WARP_F;                 // [s]  (actor input, row stride XS)
    float* xsc = xsa + 32 * XS;                           // [s; A(s)]
    float* dsm = xsc + 32 * XS;
    float* team_base = sm + n_warps * Geo::WARP_F + team * Geo::TEAM_F;
    float* tq0 = team_base;
    float* tq1 = team_base + (WC > 1 ? WC * 32 : 0);
    float* hs = team_base + (WC > 1 ? 2 * WC * 32 : 0);
    float* acc = sm + n_warps * Geo::WARP_F + teams * Geo::TEAM_F + team * F.n_x;

    UnitNet<NS, 1> A;  A.load(F.pA, F.ha, 0, lane);
    UnitNet<NIN, UPLC> C; C.template load<L2>(F.pC, F.hc, tw * UPLC * 32, lane);
    UnitGrads<NS, 1, 1> G; G.zero();
    float s_q = 0.f, s_n = 0.f, s_d = 0.f;
    const int n_tiles = (F.batch + 31) / 32;
    const int stride = gridDim.x * teams;
    const int iters = (n_tiles + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
        const int tile = blockIdx.x * teams + team + it * stride;
        if (WC == 1 && tile >= n_tiles) break;
        const int idx = tile * 32 + lane;
        const bool valid = tile < n_tiles && idx < F.batch;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const float sj = valid ? F.bs[(size_t)idx * NS + j] : 0.f;
            xsa[lane * XS + j] = sj; xsc[lane * XS + j] = sj;
        }
        __syncwarp();
        const float a = out_act(F.actA2, A.template forward<XS, false>(xsa, nullptr, 0, 0, lane) + A.b2);
        xsc[lane * XS + NS] = a;
        __syncwarp();
        const float q = team_sum<WC>(C.template forward<XS, !Geo::RECOMP>(xsc, hs, Geo::HP, tw * UPLC * 32, lane), tq0, tw, lane) + C.b2;
        if (valid) { s_q += q; s_n += 1.f; }
        // dq_i/da_i through the critic: upstream delta 1 per valid sample (the -1/n factor is applied in the tail)
        dsm[lane] = valid ? 1.f : 0.f;
        __syncwarp();
        float dqda;
        {
            UnitGrads<NIN, UPLC, 1> Gd;                    // discarded: only the input gradient is wanted
            Gd.zero();
            dqda = team_sum<WC>(unit_backward<NIN, UPLC, 1, XS, Geo::RECOMP, true, NS>(C, Gd, xsc, hs, Geo::HP, tw * UPLC * 32, lane, dsm, dsm),
                                tq1, tw, lane);
        }
        const float d = valid ? dqda * out_act_grad(F.actA2, a) : 0.f;
        __syncwarp();
        dsm[lane] = d;
        if (valid) s_d += d;
        __syncwarp();
        if (tw == 0) unit_backward<NS, 1, 1, XS, true, false, 0>(A, G, xsa, nullptr, 0, 0, lane, dsm, dsm);
        __syncwarp();
    }
    if (blockIdx.x == 0) tl_mark(F, L2 ? 5 : 1);
    if (tw == 0) G.store(acc, 0, F.ha, 0, lane);
    s_q = warp_sum(s_q); s_n = warp_sum(s_n); s_d = warp_sum(s_d);
    if (tw == 0 && lane == 0) {
        acc[F.nA - 1] = s_d;                                // b2 gradient
        float* t = acc + F.nA;
        t[0] = s_q; t[1] = s_n; t[2] = t[3] = t[4] = t[5] = t[6] = t[7] = 0.f;
    }
    __syncthreads();
    const float* acc0 = sm + n_warps * Geo::WARP_F + teams * Geo::TEAM_F;
    for (int q = threadIdx.x; q < F.n_x; q += blockDim.x) {
        float sum = 0.f;
        for (int t = 0; t < teams; ++t) sum += acc0[t * F.n_x + q];
        out[q] = sum;
    }
}

// ---- two-launch form: any grid; the last CTA to finish runs the tail (global partials + ticket) --------------------------------
template <int NS, int UPLC, int WC>
__global__ void __launch_bounds__(256) ddpg_fast_critic_kernel(const __grid_constant__ FastArgs F) {
    extern __shared__ __align__(16) float sm[];
    if (blockIdx.x == 0) tl_mark(F, 0);
    const unsigned long long offset = critic_tiles<NS, UPLC, WC>(F, sm, F.partials + (size_t)blockIdx.x * F.n_x);
    fast_tail(F, 0, offset);
}

template <int NS, int UPLC, int WC>
__global__ void __launch_bounds__(256) ddpg_fast_actor_kernel(const __grid_constant__ FastArgs F) {
    extern __shared__ __align__(16) float sm[];
    if (blockIdx.x == 0) tl_mark(F, 0);
    actor_tiles<NS, UPLC, WC, false>(F, sm, F.partials + (size_t)blockIdx.x * F.n_x);
    fast_tail(F, 1, 0);
}

// gradient assembly + ADAM + Polyak for one parameter (Flux ADAM on Float32 arrays with Float64 hyper-parameters)
__device__ __forceinline__ void adam_polyak(const FastArgs& F, int q, double Gq, double scale, double bp1, double bp2, float m, float v, float x,
                                            float tgt) {
    const float g = (float)(scale * Gq);
    F.grads[q] = g;
    const double gi = g;
    const float mi = (float)(F.b1 * (double)m + (1.0 - F.b1) * gi);
    const float vi = (float)(F.b2 * (double)v + (1.0 - F.b2) * gi * gi);
    F.m[q] = mi; F.v[q] = vi;
    const float delta = (float)((double)mi / (1.0 - bp1) / (sqrt((double)vi / (1.0 - bp2)) + F.eps) * F.eta);
    const float xn = x - delta;
    F.x[q] = xn;
    F.target[q] = F.polyak * tgt + (1.f - F.polyak) * xn;
}

// Last CTA: fixed-order reduction of the CTA partials, cross-rank exchange over NVLink peer memory, gradient assembly,
// ADAM, Polyak, statistics and losses.  phase 0 = critic (vector [G_A | G_B | tail]), 1 = actor ([G | tail]).
__device__ void fast_tail(const FastArgs& F, int phase, unsigned long long rng_offset) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(F.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    tl_mark(F, 2);
    for (int q = threadIdx.x; q < F.n_x; q += blockDim.x) F.xbuf[q] = (float)ordered_sum_cg(F.partials, (int)gridDim.x, (size_t)F.n_x, q);
    __syncthreads();
    tl_mark(F, 3);
    if (F.cm.nranks > 1) comm_allreduce_cta(F.cm, F.xbuf, F.xbuf, F.n_x);
    tl_mark(F, 4);
    const int n_acc = phase == 0 ? F.nC : F.nA;
    const float* tl = F.xbuf + (phase == 0 ? 2 * F.nC : F.nA);
    const double bp1 = F.betap[0], bp2 = F.betap[1];
    double n, rbar = 0.0, scale;
    if (phase == 0) { n = (double)tl[2]; rbar = (double)tl[0] / n; scale = -2.0 / n; }
    else { n = (double)tl[1]; scale = -1.0 / n; }
    for (int q = threadIdx.x; q < n_acc; q += blockDim.x) {
        double Gq = (double)F.xbuf[q];
        if (phase == 0 && F.literal) Gq += rbar * (double)F.xbuf[F.nC + q];
        adam_polyak(F, q, Gq, scale, bp1, bp2, F.m[q], F.v[q], F.x[q], F.target[q]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* st = F.stats;
        if (phase == 0) {
            st[ST_R] = (double)tl[0]; st[ST_R2] = (double)tl[1]; st[ST_N] = n; st[ST_C] = (double)tl[3]; st[ST_C2] = (double)tl[4]; st[ST_Q] = 0.0;
            if (F.fetch) F.dev->rng_offset = rng_offset + (unsigned long long)F.batch;
        } else {
            st[ST_Q] = (double)tl[0];
            F.losses[0] = critic_loss_from(st, F.literal);
            F.losses[1] = (float)(-(double)tl[0] / n);
        }
        F.betap[0] = bp1 * F.b1; F.betap[1] = bp2 * F.b2;
        *F.ticket = 0;
        if (F.tl) {
            const unsigned long long k = F.tl[0];
            if (k < 500) {
                unsigned long long* r = F.tl + 8 + k * 8;
                r[0] = (unsigned long long)phase;
                for (int j = 0; j < 5; ++j) r[1 + j] = F.tl[4096 + j];
                r[6] = global_timer_ns();
                F.tl[0] = k + 1;
            }
        }
    }
}

// ---- one-launch form: the whole update in ONE kernel run by a single thread-block cluster (<= 16 CTAs) --------------------------
// The CTA partials never leave the SMs: every CTA keeps its partial in shared memory, the cluster synchronises, and CTA r
// reduces slice r of the vector by reading the 16 partials through DISTRIBUTED SHARED MEMORY (fixed order: identical sums to
// the two-launch form), exchanges the slice with the peer GPUs under its own flag (16 slices in flight at once), and applies
// ADAM + Polyak to its slice with the optimiser state it prefetched at kernel entry.  A second cluster barrier publishes the
// updated critic, the actor phase follows in the same launch: no ticket, no __threadfence round trips, no launch gap.
constexpr int kSlicePF = 2;          // parameters per thread and slice (<= 2 * 256 * 16 parameters per network)

struct SliceState { float m[kSlicePF], v[kSlicePF], x[kSlicePF], t[kSlicePF]; };

__device__ __forceinline__ void slice_prefetch(const FastArgs& F, int q0, int q1, SliceState& S) {
#pragma unroll
    for (int k = 0; k < kSlicePF; ++k) {
        const int q = q0 + (int)threadIdx.x + k * (int)blockDim.x;
        const bool ok = q < q1;
        S.m[k] = ok ? F.m[q] : 0.f; S.v[k] = ok ? F.v[q] : 0.f; S.x[k] = ok ? F.x[q] : 0.f; S.t[k] = ok ? F.target[q] : 0.f;
    }
}

// reduce + exchange + apply for this CTA's slice.  accsum: this CTA's partial in shared memory (same offset in every CTA of the
// cluster); xv: shared scratch for the slice payload [G_A slice | G_B slice | 8 tail scalars].  Returns the tail in tail8 (all threads).
__device__ __forceinline__ void slice_apply(const FastArgs& F, int phase, cooperative_groups::cluster_group& cluster, float* accsum, float* xv,
                                            const SliceState& S, int q0, int q1, unsigned int epoch, double bp1, double bp2, float (&tail8)[8]) {
    const int nblk = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int n_acc = phase == 0 ? F.nC : F.nA;
    const int len = max(q1 - q0, 0);
    const int nsets = phase == 0 ? 2 : 1;
    const int tail_off = phase == 0 ? 2 * F.nC : F.nA;
    // payload index p: [0, len) set A, [len, 2 len) set B (critic), then 8 tail scalars
    const int np = nsets * len + 8;
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        int src;
        if (p < nsets * len) { const int set = p / max(len, 1), j = p - set * len; src = set * n_acc + q0 + j; }
        else src = tail_off + (p - nsets * len);
        double sum = 0.0;
        for (int r = 0; r < nblk; ++r) sum += (double)cluster.map_shared_rank(accsum, r)[src];
        xv[p] = (float)sum;
    }
    __syncthreads();
    if (F.cm.nranks > 1) comm_allreduce_slice(F.cm, epoch, rank, xv, xv, np);
#pragma unroll
    for (int j = 0; j < 8; ++j) tail8[j] = xv[nsets * len + j];
    double n, rbar = 0.0, scale;
    if (phase == 0) { n = (double)tail8[2]; rbar = (double)tail8[0] / n; scale = -2.0 / n; }
    else { n = (double)tail8[1]; scale = -1.0 / n; }
#pragma unroll
    for (int k = 0; k < kSlicePF; ++k) {
        const int j = (int)threadIdx.x + k * (int)blockDim.x;
        if (j < len) {
            double Gq = (double)xv[j];
            if (phase == 0 && F.literal) Gq += rbar * (double)xv[len + j];
            adam_polyak(F, q0 + j, Gq, scale, bp1, bp2, S.m[k], S.v[k], S.x[k], S.t[k]);
        }
    }
}

template <int NS, int UPLC, int WC>
__global__ void __launch_bounds__(256) ddpg_fast_cluster_kernel(const __grid_constant__ FastArgs Fc, const __grid_constant__ FastArgs Fa) {
    namespace cg = cooperative_groups;
    using Geo = FastGeom<NS, UPLC, WC>;
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int nblk = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int n_warps = blockDim.x >> 5;
    float* accsum = sm + Geo::base_floats(n_warps, Fc.n_x);          // [n_x of the critic phase] (the actor's vector is shorter)
    float* xv = accsum + Fc.n_x;                                      // [<= n_x] slice payload
    if (rank == 0) tl_mark(Fc, 0);
    // this CTA's slices and their optimiser state: requested now, consumed after the gradient loops
    const int slc = (Fc.nC + nblk - 1) / nblk, sla = (Fa.nA + nblk - 1) / nblk;
    const int c0 = rank * slc, c1 = min(Fc.nC, c0 + slc), a0 = rank * sla, a1 = min(Fa.nA, a0 + sla);
    SliceState Sc, Sa;
    slice_prefetch(Fc, c0, c1, Sc);
    slice_prefetch(Fa, a0, a1, Sa);
    const double cbp1 = Fc.betap[0], cbp2 = Fc.betap[1], abp1 = Fa.betap[0], abp2 = Fa.betap[1];
    const unsigned int e0 = Fc.cm.nranks > 1 ? Fc.cm.epoch[0] : 0u;
    // ---- critic phase --------------------------------------------------------------------------------------------------------
    const unsigned long long offset = critic_tiles<NS, UPLC, WC>(Fc, sm, accsum);
    cluster.sync();
    if (rank == 0) tl_mark(Fc, 2);
    float tc[8], ta[8];
    slice_apply(Fc, 0, cluster, accsum, xv, Sc, c0, c1, e0 + 1, cbp1, cbp2, tc);
    if (rank == 0) tl_mark(Fc, 3);
    cluster.sync();                       // every remote read of accsum is done; the updated critic is visible cluster-wide
    // ---- actor phase (through the UPDATED critic) ---------------------------------------------------------------------------
    actor_tiles<NS, UPLC, WC, true>(Fa, sm, accsum);
    cluster.sync();
    if (rank == 0) tl_mark(Fc, 4);
    slice_apply(Fa, 1, cluster, accsum, xv, Sa, a0, a1, e0 + 2, abp1, abp2, ta);
    cluster.sync();                       // shared memory of every CTA stays alive until the last remote read
    if (rank == 0 && threadIdx.x == 0) {
        double* st = Fc.stats;
        const double n = (double)tc[2];
        st[ST_R] = (double)tc[0]; st[ST_R2] = (double)tc[1]; st[ST_N] = n; st[ST_C] = (double)tc[3]; st[ST_C2] = (double)tc[4];
        st[ST_Q] = (double)ta[0];
        Fa.losses[0] = critic_loss_from(st, Fc.literal);
        Fa.losses[1] = (float)(-(double)ta[0] / (double)ta[1]);
        if (Fc.fetch) Fc.dev->rng_offset = offset + (unsigned long long)Fc.batch;
        Fc.betap[0] = cbp1 * Fc.b1; Fc.betap[1] = cbp2 * Fc.b2;
        Fa.betap[0] = abp1 * Fa.b1; Fa.betap[1] = abp2 * Fa.b2;
        if (Fc.cm.nranks > 1) Fc.cm.epoch[0] = e0 + 2;
        if (Fc.tl) {
            const unsigned long long k = Fc.tl[0];
            if (k < 500) {
                // {2, entry, critic loop done, after cluster.sync, critic applied, actor loop done, after cluster.sync, end}
                unsigned long long* r = Fc.tl + 8 + k * 8;
                r[0] = 2ull;
                r[1] = Fc.tl[4096]; r[2] = Fc.tl[4097]; r[3] = Fc.tl[4098]; r[4] = Fc.tl[4099]; r[5] = Fc.tl[4101]; r[6] = Fc.tl[4100];
                r[7] = global_timer_ns();
                Fc.tl[0] = k + 1;
            }
        }
    }
}
