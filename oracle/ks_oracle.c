/*
 * CPU oracle, C restatement of the reference's KS environment step.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never linked into the product.
 * Built by oracle/Makefile into oracle/_ref/libks_oracle.so and used by
 * tests/ (cross-check against oracle/ks_oracle.py, which is pinned by the
 * golden trajectories) and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Restates, operation by operation (Float64, complex FFTs on zero-imaginary data,
 * 123 FFTs per env step like the reference):
 *   do_step          /root/reference/scripts/KS/setup/KSSetup.jl:130-160
 *   prepare_action   KSSetup.jl:231-245      (dense n_act x nx loop)
 *   reward_function  KSSetup.jl:162-184      (dense dots)
 *   featurize        KSSetup.jl:190-229      (dense dots + circshift windows)
 *   env(action)      /root/reference/src/PDEenv.jl:195-241
 * The FFT is FFTW in the reference (third-party, not in the tree); here it is a
 * plain recursive mixed-radix (2,3,5) Cooley-Tukey DFT.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

typedef struct {
    int n;
    cplx* w;      /* w[k] = exp(-2 pi i k / n) */
    cplx* tmp;
} plan_t;

static void plan_init(plan_t* p, int n) {
    p->n = n;
    p->w = (cplx*)malloc(sizeof(cplx) * n);
    p->tmp = (cplx*)malloc(sizeof(cplx) * n * 2);
    for (int k = 0; k < n; ++k) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * k / n;
        p->w[k] = (double)cosl(a) + I * (double)sinl(a);
    }
}
static void plan_free(plan_t* p) { free(p->w); free(p->tmp); }

/* out[0..n) = DFT of in[0], in[s], in[2s], ...; ws = N/n stride into the root table */
static void fft_rec(const plan_t* P, int n, const cplx* in, int s, cplx* out, int sign) {
    if (n == 1) { out[0] = in[0]; return; }
    int p = (n % 4 == 0) ? 4 : (n % 2 == 0) ? 2 : (n % 3 == 0) ? 3 : (n % 5 == 0) ? 5 : n;
    int m = n / p;
    int ws = P->n / n;
    if (p == n && n > 5) {  /* generic prime: O(n^2) */
        for (int k = 0; k < n; ++k) {
            cplx acc = 0;
            for (int j = 0; j < n; ++j) {
                cplx w = P->w[((long)j * k % n) * ws];
                acc += in[j * s] * (sign < 0 ? w : conj(w));
            }
            out[k] = acc;
        }
        return;
    }
    for (int j = 0; j < p; ++j) fft_rec(P, m, in + j * s, s * p, out + j * m, sign);
    /* combine: X[k + q m] = sum_j W_n^{j k} W_p^{j q} Y_j[k] */
    for (int k = 0; k < m; ++k) {
        cplx y[5];
        for (int j = 0; j < p; ++j) {
            cplx w = P->w[(j * k) * ws];
            y[j] = out[j * m + k] * (sign < 0 ? w : conj(w));
        }
        if (p == 2) {
            out[k] = y[0] + y[1];
            out[k + m] = y[0] - y[1];
        } else if (p == 4) {
            cplx a = y[0] + y[2], b = y[0] - y[2], c = y[1] + y[3], d = (y[1] - y[3]) * (sign < 0 ? -I : I);
            out[k] = a + c; out[k + m] = b + d; out[k + 2 * m] = a - c; out[k + 3 * m] = b - d;
        } else {
            for (int q = 0; q < p; ++q) {
                cplx acc = y[0];
                for (int j = 1; j < p; ++j) {
                    cplx w = P->w[((j * q) % p) * (P->n / p)];
                    acc += y[j] * (sign < 0 ? w : conj(w));
                }
                P->tmp[q] = acc;
            }
            for (int q = 0; q < p; ++q) out[k + q * m] = P->tmp[q];
        }
    }
}

static void fft(const plan_t* P, cplx* x, int sign) {   /* in place via scratch */
    cplx* t = P->tmp + 8;
    fft_rec(P, P->n, x, 1, t, sign);
    if (sign < 0) memcpy(x, t, sizeof(cplx) * P->n);
    else { double inv = 1.0 / P->n; for (int i = 0; i < P->n; ++i) x[i] = t[i] * inv; }
}

/* KSSetup.jl:130-160, literal (including the two redundant FFTs inside the loop). */
static void ks_do_step(const plan_t* P, int nx, double Lx, double dt, int S, double mu, const double* y,
                       const double* p, double* y_out, cplx* work) {
    cplx* u = work; cplx* Nn = u + nx; cplx* Nn1 = Nn + nx; cplx* ph = Nn1 + nx; cplx* mh = ph + nx;
    double* L = (double*)(mh + nx); double* Ainv = L + nx; double* B = Ainv + nx; double* al = B + nx;
    const double dto = dt / S, dt2 = dto / 2, dt32 = 3 * dto / 2, dx = Lx / nx;
    for (int i = 0; i < nx; ++i) {
        double kx = (i < nx / 2) ? i : (i == nx / 2 ? 0 : i - nx);
        al[i] = 2 * M_PI * kx / Lx;
        L[i] = al[i] * al[i] - al[i] * al[i] * al[i] * al[i];
        Ainv[i] = 1.0 / (1.0 - dt2 * L[i]);
        B[i] = 1.0 + dt2 * L[i];
    }
    for (int i = 0; i < nx; ++i) { u[i] = y[i]; Nn[i] = y[i] * y[i]; }
    fft(P, Nn, -1);
    for (int i = 0; i < nx; ++i) { Nn[i] *= -0.5 * I * al[i]; Nn1[i] = Nn[i]; }
    fft(P, u, -1);
    for (int n = 0; n < S; ++n) {
        for (int i = 0; i < nx; ++i) { Nn1[i] = Nn[i]; Nn[i] = u[i]; }
        fft(P, Nn, +1);
        for (int i = 0; i < nx; ++i) Nn[i] = Nn[i] * Nn[i];
        fft(P, Nn, -1);
        for (int i = 0; i < nx; ++i) Nn[i] *= -0.5 * I * al[i];
        for (int i = 0; i < nx; ++i) { ph[i] = p[i]; mh[i] = mu * cos(2 + M_PI + (dx * (i + 1)) / (Lx / 2)); }
        fft(P, ph, -1);
        fft(P, mh, -1);
        for (int i = 0; i < nx; ++i)
            u[i] = Ainv[i] * (B[i] * u[i] + dt32 * Nn[i] - dt2 * Nn1[i] + dto * ph[i]) + dto * mh[i];
    }
    fft(P, u, +1);
    for (int i = 0; i < nx; ++i) y_out[i] = creal(u[i]);
}

/* public: one do_step on one environment */
void ks_oracle_do_step(int nx, double Lx, double dt, int S, double mu, const double* y, const double* p, double* y_out) {
    plan_t P; plan_init(&P, nx);
    cplx* work = (cplx*)malloc(sizeof(cplx) * nx * 8);
    ks_do_step(&P, nx, Lx, dt, S, mu, y, p, y_out, work);
    free(work); plan_free(&P);
}

/*
 * public: `n_steps` full env steps (PDEenv.jl:195-241 with the KS closures, window W) on `n_envs`
 * environments, one after the other on the calling thread (the reference is single-threaded;
 * bench.py runs one call per host core on disjoint slices).  Arrays are env-major:
 *   y [B][nx] in/out, action_prev [B][n_a] in/out, actions [n_steps][B][n_a],
 *   g_sens [n_s][nx], g_act [n_a][nx], a2s [n_a] (0-based), state_out [B][n_a][W], reward_out [B][n_a].
 */
void ks_oracle_env_steps(int nx, double Lx, double dt, int S, double mu, int n_envs, int n_steps, int n_s, int n_a,
                         int W, double agent_power, double max_value, double a_pun, double da_pun,
                         const double* g_sens, const double* g_act, const int* a2s, double* y, double* action_prev,
                         const double* actions, double* state_out, double* reward_out, int n_threads) {
    (void)n_threads;   /* single-threaded like the reference; callers parallelise over env slices */
    {
        plan_t P; plan_init(&P, nx);
        cplx* work = (cplx*)malloc(sizeof(cplx) * nx * 8);
        double* p = (double*)malloc(sizeof(double) * nx);
        double* yn = (double*)malloc(sizeof(double) * nx);
        double* sens = (double*)malloc(sizeof(double) * n_s);
        for (int b = 0; b < n_envs; ++b) {
            double* yb = y + (size_t)b * nx;
            double* ap = action_prev + (size_t)b * n_a;
            for (int st = 0; st < n_steps; ++st) {
                const double* a = actions + ((size_t)st * n_envs + b) * n_a;
                /* prepare_action, KSSetup.jl:238-242 */
                for (int i = 0; i < nx; ++i) p[i] = 0.0;
                for (int j = 0; j < n_a; ++j) {
                    const double c = agent_power * a[j];
                    const double* g = g_act + (size_t)j * nx;
                    for (int i = 0; i < nx; ++i) p[i] = p[i] + c * g[i];
                }
                ks_do_step(&P, nx, Lx, dt, S, mu, yb, p, yn, work);
                memcpy(yb, yn, sizeof(double) * nx);
                /* reward_function, KSSetup.jl:162-178 */
                for (int j = 0; j < n_a; ++j) {
                    const double* g = g_sens + (size_t)a2s[j] * nx;
                    double d = 0.0;
                    for (int i = 0; i < nx; ++i) d += (yb[i] * 6) * g[i];
                    const double s = pow(fabs(d), 1.3) / (max_value * 3);
                    const double da = a[j] - ap[j];
                    reward_out[(size_t)b * n_a + j] = -fabs(s) - a_pun * a[j] * a[j] - da_pun * da * da;
                }
                /* featurize, KSSetup.jl:197-207 */
                for (int i = 0; i < n_s; ++i) {
                    const double* g = g_sens + (size_t)i * nx;
                    double d = 0.0;
                    for (int k = 0; k < nx; ++k) d += yb[k] * g[k];
                    sens[i] = d / max_value;
                }
                const int h = W / 2;
                for (int j = 0; j < n_a; ++j)
                    for (int r = 0; r < W; ++r) {
                        int idx = (a2s[j] - (r - h)) % n_s;
                        if (idx < 0) idx += n_s;
                        state_out[((size_t)b * n_a + j) * W + r] = sens[idx];
                    }
                memcpy(ap, a, sizeof(double) * n_a);
            }
        }
        free(sens); free(yn); free(p); free(work); plan_free(&P);
    }
}
