// KS back-end: spectral constant tables + launch dispatch for ks_step_kernel.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "ctx.hpp"
#include "ks_step.cuh"
#include "ks_step_tm.cuh"

namespace pdeb200 {

namespace {

struct Fact { int n, n1, n2; };
// N = N1*N2 four-step factorizations with generated in-register DFTs for both factors.
const Fact kFacts[] = {{64, 8, 8},     {128, 8, 16},   {192, 12, 16}, {240, 15, 16},
                       {256, 16, 16},  {320, 16, 20},  {384, 16, 24}, {512, 16, 32},
                       {600, 24, 25},  {1024, 32, 32}};

template <typename T>
int32_t upload(pdeb200_ctx* c, void** dst, const std::vector<T>& v) {
    PDEB_CUDA(c, cudaMalloc(dst, v.size() * sizeof(T)));
    PDEB_CUDA(c, cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PDEB200_OK;
}

template <typename T>
int32_t setup_t(pdeb200_ctx* c) {
    using C = typename V2<T>::type;
    const pdeb200_config& g = c->cfg;
    const int N = g.nx, N1 = c->N1, N2 = c->N2;
    const long double two_pi = 6.283185307179586476925286766559L;
    std::vector<C> t12(N), t21(N);
    for (int k = 0; k < N1; ++k)
        for (int t = 0; t < N2; ++t) {
            long double a = -two_pi * (long double)((long long)k * t % N) / N;
            t12[k * N2 + t] = V2<T>::make((T)cosl(a), (T)sinl(a));
        }
    for (int k = 0; k < N2; ++k)
        for (int t = 0; t < N1; ++t) {
            long double a = -two_pi * (long double)((long long)k * t % N) / N;
            t21[k * N1 + t] = V2<T>::make((T)cosl(a), (T)sinl(a));
        }
    // KSSetup.jl:115-119 and :131-135, evaluated in Float64 like the reference.
    const double h = g.dt / g.oversampling, dt2 = h / 2;
    std::vector<T> c1(N), cN(N), ah(N);
    for (int i = 0; i < N; ++i) {
        double kx = (i < N / 2) ? i : (i == N / 2 ? 0 : i - N);
        double alpha = 2 * M_PI * kx / g.Lx;
        double L = alpha * alpha - alpha * alpha * alpha * alpha;
        double Ainv = 1.0 / (1.0 - dt2 * L);
        double B = 1.0 + dt2 * L;
        c1[i] = (T)(Ainv * B);
        cN[i] = (T)(Ainv * (-0.5 * alpha) / ((double)N * (double)N) * (3 * h / 2));   // the 3h/2 of AB2 folded in
        ah[i] = (T)(Ainv * h);
    }
    int32_t rc;
    if ((rc = upload<C>(c, &c->tw12, t12))) return rc;
    if ((rc = upload<C>(c, &c->tw21, t21))) return rc;
    if ((rc = upload<T>(c, &c->c1, c1))) return rc;
    if ((rc = upload<T>(c, &c->cN, cN))) return rc;
    if ((rc = upload<T>(c, &c->ainvh, ah))) return rc;
    if (g.mu != 0.0) {
        // h * fft(mu * cos(2 + pi + x/(Lx/2))), KSSetup.jl:155 (quirk Q3), naive DFT in long double
        std::vector<long double> f(N);
        const double dx = g.Lx / N;
        for (int n = 0; n < N; ++n) f[n] = g.mu * std::cos(2 + M_PI + (dx * (n + 1)) / (g.Lx / 2));
        std::vector<C> hm(N);
        for (int k = 0; k < N; ++k) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; ++n) {
                long double a = -two_pi * (long double)((long long)k * n % N) / N;
                sr += f[n] * cosl(a); si += f[n] * sinl(a);
            }
            hm[k] = V2<T>::make((T)(h * sr), (T)(h * si));
        }
        if ((rc = upload<C>(c, &c->hm, hm))) return rc;
    }
    return PDEB200_OK;
}

// Pick the CTA size (warps) and residency so that the grid fills whole waves: with deterministic,
// equal work per pair the step time is (rounds) x (time of one resident set), so a 1.73-wave grid
// costs as much as a 2.0-wave one.  cost ~ rounds * resident_pairs; ties go to more resident warps.
template <typename T, int N1, int N2>
void plan(const pdeb200_ctx* c, int n_sm, int* warps_out, int* ctas_per_sm_out, bool lowreg) {
    using G = KsGeom<N1, N2>;
    constexpr int ppw = 32 / G::TP;
    const int max_warps = lowreg ? KsMaxWarps<T, true>::value : KsMaxWarps<T, false>::value;
    const int n_pairs = (c->cfg.n_envs + 1) / 2;
    static const int forced = [] { const char* e = getenv("PDEB200_KS_WARPS"); return e ? atoi(e) : 0; }();
    double best = 1e300; int bw = 1, bc = 1;
    for (int w = 1; w <= max_warps; ++w) {
        if (forced && w != forced) continue;
        const size_t smem = ks_smem_bytes<T, N1, N2>(w * ppw, c->cfg.oversampling > 1) + 1024;
        int cps = std::min((int)((227 * 1024) / smem), max_warps / w);
        if (cps < 1) continue;
        const int ppc = w * ppw;
        const int n_ctas = (n_pairs + ppc - 1) / ppc;
        const int per_sm = (n_ctas + n_sm - 1) / n_sm;
        cps = std::min(cps, per_sm);
        const int rounds = (per_sm + cps - 1) / cps;
        // more resident warps hide latency better: mild bonus
        const double cost = (double)rounds * cps * ppc * (1.0 + 0.02 * (max_warps - cps * w));
        if (cost < best - 1e-9) { best = cost; bw = w; bc = cps; }
    }
    *warps_out = bw; *ctas_per_sm_out = bc;
}

// Layout of the natural-order state copy the sensor gather reads (see the epilogue of ks_step_kernel): if the
// sensors' first taps are equally spaced by `sp` grid points (sp | N), point n goes to plane n % sp, slot n / sp,
// planes padded by one entry against bank conflicts on the write side; the table's indices are mapped once here.
template <int N1, int N2>
int32_t sensor_layout(pdeb200_ctx* c) {
    using G = KsGeom<N1, N2>;
    constexpr int N = G::N;
    c->ks_layout_dirty = false; c->ks_layout_on = false;
    static const bool off = [] { const char* e = getenv("PDEB200_KS_PLAIN_SENSOR_LAYOUT"); return e && atoi(e) != 0; }();
    const int ns = c->sens.n_rows, nnz = c->sens.nnz_max;
    if (off || ns < 2) return PDEB200_OK;
    std::vector<int> idx((size_t)nnz * ns);
    PDEB_CUDA(c, cudaMemcpy(idx.data(), c->sens.d_idx, idx.size() * sizeof(int), cudaMemcpyDeviceToHost));
    // most frequent distance between the first taps of neighbouring sensors (windows that wrap around the periodic
    // boundary list their taps in a different order; any permutation is CORRECT, the spacing only decides how well
    // the gather coalesces)
    std::vector<int> votes(N, 0);
    for (int i = 0; i + 1 < ns; ++i) votes[((idx[i + 1] - idx[i]) % N + N) % N]++;
    const int sp = (int)(std::max_element(votes.begin(), votes.end()) - votes.begin());
    if (sp < 2 || N % sp != 0 || 2 * votes[sp] < ns) return PDEB200_OK;
    const int slots = N / sp;
    const int stride = (sp * (slots + 1) <= G::XB) ? slots + 1 : slots;
    std::vector<int> perm(N);
    for (int n = 0; n < N; ++n) perm[n] = (n % sp) * stride + n / sp;
    for (int& v : idx) v = perm[v];
    if (!c->ks_perm) PDEB_CUDA(c, cudaMalloc(&c->ks_perm, N * sizeof(int)));
    if (c->ks_sens_idx) { cudaFree(c->ks_sens_idx); c->ks_sens_idx = nullptr; }
    PDEB_CUDA(c, cudaMalloc(&c->ks_sens_idx, idx.size() * sizeof(int)));
    PDEB_CUDA(c, cudaMemcpy(c->ks_perm, perm.data(), N * sizeof(int), cudaMemcpyHostToDevice));
    PDEB_CUDA(c, cudaMemcpy(c->ks_sens_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->ks_layout_on = true;
    return PDEB200_OK;
}

// fp64 N = 256: the tensor-memory variant (ks_step_tm.cuh).  One CTA of up to 8 warps per SM (512 TMEM columns), or two CTAs of
// up to 4 warps (256 columns each); the CTA size is chosen like in plan() so that the grid fills whole waves.
int32_t launch_tm(pdeb200_ctx* c, const KsArgs<double>& A, int n_sm, bool spec) {
    const int n_pairs = (c->cfg.n_envs + 1) / 2;
    static const int forced = [] { const char* e = getenv("PDEB200_KS_WARPS"); return e ? atoi(e) : 0; }();
    // A CTA of up to 4 warps holds 256 of the SM's 512 TMEM columns whatever its size, so 4-warp CTAs are the smallest that
    // let two launches of different env shards (bench.py's e2e leg) fill an SM together: smaller ones only for tiny batches.
    double best = 1e300; int bw = std::min(4, (n_pairs + 1) / 2);
    for (int w = 8; w >= 4; --w) {
        if (n_pairs < 8 && !forced) break;
        if (forced && w != forced) continue;
        const int ppc = 2 * w;
        const int n_ctas = (n_pairs + ppc - 1) / ppc;
        const int per_sm = (n_ctas + n_sm - 1) / n_sm;
        const int cps = std::min(w <= 4 ? 2 : 1, per_sm);
        const int rounds = (per_sm + cps - 1) / cps;
        const double cost = (double)rounds * cps * ppc * (1.0 + 0.02 * (8 - cps * w));
        if (cost < best - 1e-9) { best = cost; bw = w; }
    }
    if (forced && forced >= 1 && forced <= 8) bw = forced;
    const int PAIRS = 2 * bw;
    const size_t n_tab = spec ? 0 : (size_t)A.sens.nnz_max * A.n_sensors;
    const size_t smem = ks_tm_smem_bytes(PAIRS, n_tab);
    if (smem > 227 * 1024) return PDEB200_EUNSUPPORTED;
    auto kern = spec ? ks_step_tm_kernel<true> : ks_step_tm_kernel<false>;
    c->core_kernel = spec ? "ks_step_tm_kernel<spectral sensors>" : "ks_step_tm_kernel<gather sensors>";
    PDEB_CUDA(c, ensure_dyn_smem(kern, smem, c->device));
    kern<<<(n_pairs + PAIRS - 1) / PAIRS, bw * 32, smem, c->stream>>>(A);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}
template <typename T> int32_t launch_tm(pdeb200_ctx*, const KsArgs<T>&, int, bool) { return PDEB200_EUNSUPPORTED; }

template <typename T, int N1, int N2>
int32_t launch(pdeb200_ctx* c) {
    using G = KsGeom<N1, N2>;
    using C = typename V2<T>::type;
    const pdeb200_config& g = c->cfg;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
    int warps, cps;
    // few substeps: the low-register variant (12 warps/SM in fp64); PDEB200_KS_LOWREG=0/1 overrides for experiments
    static const int force_low = [] { const char* e = getenv("PDEB200_KS_LOWREG"); return e ? atoi(e) : -1; }();
    const bool lowreg = sizeof(T) == 8 && G::RMAX <= 16 && (force_low >= 0 ? force_low != 0 : g.oversampling <= 4);
    plan<T, N1, N2>(c, n_sm, &warps, &cps, lowreg);
    const int PAIRS = warps * 32 / G::TP;
    KsArgs<T> A;
    A.n_envs = g.n_envs; A.S = g.oversampling; A.n_sensors = g.n_sensors;
    A.tw12 = (const C*)c->tw12; A.tw21 = (const C*)c->tw21;
    A.c1 = (const T*)c->c1; A.cN = (const T*)c->cN; A.ainvh = (const T*)c->ainvh; A.hm = (const C*)c->hm;
    const double h = g.dt / g.oversampling;
    A.third = (T)((h / 2) / (3 * h / 2)); A.inv_n = (T)(1.0 / G::N);
    if (c->ks_layout_dirty) {
        int32_t rc = sensor_layout<N1, N2>(c);
        if (rc) return rc;
    }
    A.sens = EllTable<T>{c->ks_layout_on ? c->ks_sens_idx : c->sens.d_idx, (const T*)c->sens.d_w, c->sens.nnz_max, c->sens.n_rows};
    A.perm = c->ks_layout_on ? c->ks_perm : nullptr;
    A.sens_hat = (const C*)c->ks_sens_hat; A.sens_sp = c->ks_sens_sp;
    A.y = (T*)c->y; A.p = (const T*)c->p; A.sensors_out = (T*)c->sensors; A.vmax_out = (T*)c->vmax;
    const int n_pairs = (g.n_envs + 1) / 2;
    const int grid = (n_pairs + PAIRS - 1) / PAIRS;
    const size_t smem = ks_smem_bytes<T, N1, N2>(PAIRS, g.oversampling > 1);
    // spectral sensor dots where they pay (few substeps); PDEB200_KS_SPECTRAL_SENSORS=2 forces them at any oversampling
    static const int spec_mode = [] { const char* e = getenv("PDEB200_KS_SPECTRAL_SENSORS"); return e ? atoi(e) : 1; }();
    // tensor-memory variant: fp64, 16 x 16, three or more substeps (measured: 165 vs 194 us at oversampling 4 and 32768 envs, 410 vs
    // 386 us at oversampling 1 and 131072 envs; PDEB200_KS_TMEM=0 off, 2 = at any oversampling)
    static const int tm_mode = [] { const char* e = getenv("PDEB200_KS_TMEM"); return e ? atoi(e) : 1; }();
    if (sizeof(T) == 8 && N1 == 16 && N2 == 16 && tm_mode && (tm_mode == 2 || g.oversampling > 2)) {
        const int32_t rc = launch_tm(c, A, n_sm, c->ks_sens_hat != nullptr && spec_mode != 0);
        if (rc != PDEB200_EUNSUPPORTED) return rc;
    }
    const bool spec = c->ks_sens_hat != nullptr && (spec_mode == 2 || g.oversampling <= 4);
    auto kern = lowreg ? (spec ? ks_step_kernel<T, N1, N2, true, true> : ks_step_kernel<T, N1, N2, true, false>)
                       : (spec ? ks_step_kernel<T, N1, N2, false, true> : ks_step_kernel<T, N1, N2, false, false>);
    c->core_kernel = sizeof(T) == 8 ? (lowreg ? "ks_step_kernel<f64, low-register>" : "ks_step_kernel<f64>") : "ks_step_kernel<f32>";
    PDEB_CUDA(c, ensure_dyn_smem(kern, smem, c->device));
    kern<<<grid, warps * 32, smem, c->stream>>>(A);
    PDEB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return PDEB200_OK;
}

template <typename T>
int32_t dispatch(pdeb200_ctx* c) {
    switch (c->cfg.nx) {
        case 64:   return launch<T, 8, 8>(c);
        case 128:  return launch<T, 8, 16>(c);
        case 192:  return launch<T, 12, 16>(c);
        case 240:  return launch<T, 15, 16>(c);
        case 256:  return launch<T, 16, 16>(c);
        case 320:  return launch<T, 16, 20>(c);
        case 384:  return launch<T, 16, 24>(c);
        case 512:  return launch<T, 16, 32>(c);
        case 600:  return launch<T, 24, 25>(c);
        case 1024: return launch<T, 32, 32>(c);
    }
    return fail(c, PDEB200_EUNSUPPORTED, "KS: unsupported nx");
}

}  // namespace

int32_t ks_setup(pdeb200_ctx* c) {
    c->N1 = 0;
    for (const Fact& f : kFacts)
        if (f.n == c->cfg.nx) { c->N1 = f.n1; c->N2 = f.n2; }
    if (!c->N1)
        return fail(c, PDEB200_EUNSUPPORTED,
                    "KS: nx must be one of 64,128,192,240,256,320,384,512,600,1024 (four-step FFT factor table)");
    if (c->cfg.oversampling < 1) return fail(c, PDEB200_EINVAL, "KS: oversampling must be >= 1");
    return c->cfg.dtype == PDEB200_F64 ? setup_t<double>(c) : setup_t<float>(c);
}

// Equally spaced, shift-invariant sensor bases (every shipped KS script: periodic Gaussians at every k-th grid point,
// KSSetup.jl:82-113): <y, g_i> = s(sp * i) with s = ifft(fft(y) . conj(fft(g_0))) -- ONE more inverse transform of the spectrum
// the kernel already holds, instead of a 21-tap gather per sensor through shared memory (a third of the kernel's shared-memory
// wavefronts at oversampling 1, profiles/r2_ks_S1.md).  Detected here from the dense basis the host hands to set_bases; rows that
// are not shifted copies of row 0 to 1e-14 (or PDEB200_KS_SPECTRAL_SENSORS=0) keep the gather.
int32_t ks_bases_changed(pdeb200_ctx* c, const double* sb) {
    const int N = c->cfg.nx, ns = c->cfg.n_sensors;
    if (c->ks_sens_hat) { cudaFree(c->ks_sens_hat); c->ks_sens_hat = nullptr; }
    c->ks_sens_sp = 0;
    static const bool off = [] { const char* e = getenv("PDEB200_KS_SPECTRAL_SENSORS"); return e && atoi(e) == 0; }();
    if (off || ns < 2 || !sb) return PDEB200_OK;
    auto peak = [&](int i) {
        int best = 0;
        for (int n = 1; n < N; ++n) if (sb[(size_t)i * N + n] > sb[(size_t)i * N + best]) best = n;
        return best;
    };
    const int p0 = peak(0), sp = ((peak(1) - p0) % N + N) % N;
    if (sp < 1 || (long long)sp * ns > N) return PDEB200_OK;
    double gmax = 0;
    for (int n = 0; n < N; ++n) gmax = std::max(gmax, std::fabs(sb[n]));
    for (int i = 1; i < ns; ++i)
        for (int n = 0; n < N; ++n)
            if (std::fabs(sb[(size_t)i * N + n] - sb[(((n - sp * i) % N) + N) % N]) > 1e-14 * gmax) return PDEB200_OK;
    // H[k] = conj(sum_n g_0[n] exp(-2 pi i n k / N)), naive DFT in long double
    const long double two_pi = 6.283185307179586476925286766559L;
    std::vector<double> hr(N), hi(N);
    for (int k = 0; k < N; ++k) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            const long double a = -two_pi * (long double)((long long)k * n % N) / N;
            sr += (long double)sb[n] * cosl(a); si += (long double)sb[n] * sinl(a);
        }
        hr[k] = (double)sr; hi[k] = (double)(-si);
    }
    if (c->cfg.dtype == PDEB200_F64) {
        std::vector<double2> h(N);
        for (int k = 0; k < N; ++k) h[k] = make_double2(hr[k], hi[k]);
        PDEB_CUDA(c, cudaMalloc(&c->ks_sens_hat, N * sizeof(double2)));
        PDEB_CUDA(c, cudaMemcpy(c->ks_sens_hat, h.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
    } else {
        std::vector<float2> h(N);
        for (int k = 0; k < N; ++k) h[k] = make_float2((float)hr[k], (float)hi[k]);
        PDEB_CUDA(c, cudaMalloc(&c->ks_sens_hat, N * sizeof(float2)));
        PDEB_CUDA(c, cudaMemcpy(c->ks_sens_hat, h.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
    }
    c->ks_sens_sp = sp;
    return PDEB200_OK;
}

int32_t ks_core(pdeb200_ctx* c) { return c->cfg.dtype == PDEB200_F64 ? dispatch<double>(c) : dispatch<float>(c); }

// Algorithmic cost of one env step (SURVEY.md 8d / DESIGN.md):
//   bytes: y in + y out + action in + obs out + reward out + done
//   flops: (2S+4) real-data FFTs at 2.5 N log2 N + pointwise + sparse bases
int32_t ks_cost(const pdeb200_ctx* c, double* bytes, double* flops) {
    const pdeb200_config& g = c->cfg;
    const double N = g.nx, S = g.oversampling, w = (double)c->esz;
    if (bytes) *bytes = 2 * N * w + g.n_actuators * (w * c->a_rows + w * c->obs_rows + w) + 1;
    if (flops) {
        const double fft = 2.5 * N * std::log2(N);
        *flops = (2 * S + 4) * fft + S * 24 * (N / 2 + 1) +
                 2.0 * (c->sens.nnz_max * (double)g.n_sensors + c->actT.nnz_max * N);
    }
    return PDEB200_OK;
}

void ks_free(pdeb200_ctx* c) {
    for (void** p : {&c->tw12, &c->tw21, &c->c1, &c->cN, &c->ainvh, &c->hm, (void**)&c->ks_perm, (void**)&c->ks_sens_idx, &c->ks_sens_hat}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
}

}  // namespace pdeb200
