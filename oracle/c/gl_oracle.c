/* C restatement of the Griffin-Lim vocoder path of xd-tts -- TEST INFRASTRUCTURE.
 *
 * PARITY UNPINNED (see oracle/__init__.py): the reference keeps this arithmetic
 * in the un-vendored crate griffin-lim 0.2.0 @ e6415314 (Cargo.lock:666-680) on
 * realfft 3.3.0 / rustfft 6.2.0 (Cargo.lock:1449,1543) with rayon-parallel
 * ndarray loops (Cargo.lock:993-1002); there is no golden vector.  This file
 * restates the librosa-0.9.2 algorithm the crate ports (slides/vocoding.typ:50)
 * for the call sites src/tacotron2/mod.rs:453-456 and src/lib.rs:141-155, and is
 * pinned against oracle/gl_oracle.py (itself checked against torch.stft /
 * torch.istft) in tests/test_oracle.py.
 *
 * It exists to be (a) a second, independent checker and (b) the timed CPU arm
 * of bench.py (cpu_baseline.kind = "port", `--impl reference`): OpenMP over
 * frames stands in for the crate's rayon loops.  Never linked into the product.
 *
 * Build: make -C oracle   ->  oracle/_build/libgl_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float re, im; } cpx;

/* ---------------------------------------------------------------- FFT plan */
typedef struct {
    int n;        /* real length (power of two) */
    int m;        /* complex length n/2 */
    cpx *tw;      /* W_m^j, j < m   (forward, e^{-2 pi i j/m}) */
    cpx *rtw;     /* W_n^k, k <= m/2 ... full m for simplicity */
    float *win;   /* periodic Hann */
} plan_t;

static plan_t *plan_new(int n) {
    plan_t *p = (plan_t *)calloc(1, sizeof(plan_t));
    p->n = n; p->m = n / 2;
    p->tw = (cpx *)malloc(sizeof(cpx) * p->m);
    p->rtw = (cpx *)malloc(sizeof(cpx) * (p->m + 1));
    p->win = (float *)malloc(sizeof(float) * n);
    for (int j = 0; j < p->m; j++) {
        double a = -2.0 * M_PI * j / p->m;
        p->tw[j].re = (float)cos(a); p->tw[j].im = (float)sin(a);
    }
    for (int k = 0; k <= p->m; k++) {
        double a = -2.0 * M_PI * k / n;
        p->rtw[k].re = (float)cos(a); p->rtw[k].im = (float)sin(a);
    }
    for (int i = 0; i < n; i++) p->win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / n));
    return p;
}
static void plan_free(plan_t *p) { free(p->tw); free(p->rtw); free(p->win); free(p); }

/* Stockham autosort radix-2 passes, out-of-place ping-pong; sign=+1 conjugates
 * the twiddles (unnormalised inverse).  Result ends up in x. */
static void cfft(const plan_t *p, cpx *x, cpx *y, int inverse) {
    const int m = p->m;
    int l = m / 2, s = 1; /* l = half-size count of butterflies per group */
    cpx *a = x, *b = y;
    /* radix-4 while possible, finish with radix-2 */
    int n_left = m;
    while (n_left >= 4) {
        const int q = n_left / 4; /* butterflies per stride block */
        for (int j = 0; j < q; j++) {
            cpx w1 = p->tw[j * s], w2 = p->tw[2 * j * s], w3 = p->tw[3 * j * s];   /* 3 j s < 3 m / 4: no wrap */
            if (inverse) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
            for (int k = 0; k < s; k++) {
                const cpx c0 = a[k + s * (j)], c1 = a[k + s * (j + q)];
                const cpx c2 = a[k + s * (j + 2 * q)], c3 = a[k + s * (j + 3 * q)];
                cpx t0 = {c0.re + c2.re, c0.im + c2.im}, t1 = {c0.re - c2.re, c0.im - c2.im};
                cpx t2 = {c1.re + c3.re, c1.im + c3.im}, t3;
                if (!inverse) { t3.re = c1.im - c3.im; t3.im = -(c1.re - c3.re); }  /* -i*(c1-c3) */
                else          { t3.re = -(c1.im - c3.im); t3.im = c1.re - c3.re; }  /* +i*(c1-c3) */
                cpx u0 = {t0.re + t2.re, t0.im + t2.im};
                cpx u1 = {t1.re + t3.re, t1.im + t3.im};
                cpx u2 = {t0.re - t2.re, t0.im - t2.im};
                cpx u3 = {t1.re - t3.re, t1.im - t3.im};
                cpx *o = b + k + s * (4 * j);
                o[0] = u0;
                o[s].re = u1.re * w1.re - u1.im * w1.im;     o[s].im = u1.re * w1.im + u1.im * w1.re;
                o[2 * s].re = u2.re * w2.re - u2.im * w2.im; o[2 * s].im = u2.re * w2.im + u2.im * w2.re;
                o[3 * s].re = u3.re * w3.re - u3.im * w3.im; o[3 * s].im = u3.re * w3.im + u3.im * w3.re;
            }
        }
        n_left /= 4; s *= 4;
        cpx *t = a; a = b; b = t;
    }
    if (n_left == 2) {
        l = 1;
        for (int k = 0; k < s; k++) {
            const cpx c0 = a[k], c1 = a[k + s];
            b[k].re = c0.re + c1.re; b[k].im = c0.im + c1.im;
            b[k + s].re = c0.re - c1.re; b[k + s].im = c0.im - c1.im;
        }
        cpx *t = a; a = b; b = t;
    }
    (void)l;
    if (a != x) memcpy(x, a, sizeof(cpx) * m);
}

/* forward real FFT of n samples (already windowed) -> K = n/2+1 bins */
static void rfft_frame(const plan_t *p, const float *x, cpx *out, cpx *w0, cpx *w1) {
    const int m = p->m;
    for (int j = 0; j < m; j++) { w0[j].re = x[2 * j]; w0[j].im = x[2 * j + 1]; }
    cfft(p, w0, w1, 0);
    out[0].re = w0[0].re + w0[0].im; out[0].im = 0.f;
    out[m].re = w0[0].re - w0[0].im; out[m].im = 0.f;
    for (int k = 1; k < m; k++) {
        const cpx a = w0[k], b = w0[m - k];
        const float er = 0.5f * (a.re + b.re), ei = 0.5f * (a.im - b.im);
        const float or_ = 0.5f * (a.im + b.im), oi = 0.5f * (b.re - a.re);
        const cpx w = p->rtw[k];
        out[k].re = er + (or_ * w.re - oi * w.im);
        out[k].im = ei + (or_ * w.im + oi * w.re);
    }
}

/* inverse real FFT: K bins -> n samples, scaled by 1/n; imag of DC/Nyquist ignored */
static void irfft_frame(const plan_t *p, const cpx *in, float *x, cpx *w0, cpx *w1) {
    const int m = p->m, n = p->n;
    w0[0].re = in[0].re + in[m].re; w0[0].im = in[0].re - in[m].re;
    for (int k = 1; k < m; k++) {
        const cpx a = in[k], b = in[m - k];
        const float er = a.re + b.re, ei = a.im - b.im;
        const float dr = a.re - b.re, di = a.im + b.im;
        const cpx w = p->rtw[k]; /* conj(w) = e^{+...} */
        const float or_ = dr * w.re + di * w.im, oi = di * w.re - dr * w.im;
        w0[k].re = er - oi; w0[k].im = ei + or_;
    }
    cfft(p, w0, w1, 1);
    const float sc = 1.0f / (float)n;
    for (int j = 0; j < m; j++) { x[2 * j] = w0[j].re * sc; x[2 * j + 1] = w0[j].im * sc; }
}

/* ------------------------------------------------------ STFT / ISTFT (frame-major) */
static inline int reflect_idx(int i, int len) { /* numpy 'reflect' for |pad| < len */
    if (i < 0) i = -i;
    if (i >= len) i = 2 * (len - 1) - i;
    return i;
}

/* spec: [T][K] */
static void stft_fm(const plan_t *p, const float *y, int len, int hop, int T, cpx *spec, int pad_constant) {
    const int n = p->n, K = p->m + 1, half = n / 2;
#pragma omp parallel
    {
        float *fr = (float *)malloc(sizeof(float) * n);
        cpx *w0 = (cpx *)malloc(sizeof(cpx) * p->m), *w1 = (cpx *)malloc(sizeof(cpx) * p->m);
#pragma omp for schedule(static)
        for (int t = 0; t < T; t++) {
            const int start = t * hop - half;
            if (start >= 0 && start + n <= len) {
                for (int i = 0; i < n; i++) fr[i] = y[start + i] * p->win[i];
            } else {
                for (int i = 0; i < n; i++) {
                    const int j = start + i;
                    float v;
                    if (j >= 0 && j < len) v = y[j];
                    else v = pad_constant ? 0.f : y[reflect_idx(j, len)];
                    fr[i] = v * p->win[i];
                }
            }
            rfft_frame(p, fr, spec + (size_t)t * K, w0, w1);
        }
        free(fr); free(w0); free(w1);
    }
}

/* spec [T][K] -> y[hop*(T-1)]; frames scratch [T][n]; ascending-frame overlap-add */
static void istft_fm(const plan_t *p, const cpx *spec, int hop, int T, float *y, float *frames, const float *wss) {
    const int n = p->n, K = p->m + 1, half = n / 2, len = hop * (T - 1);
#pragma omp parallel
    {
        cpx *w0 = (cpx *)malloc(sizeof(cpx) * p->m), *w1 = (cpx *)malloc(sizeof(cpx) * p->m);
#pragma omp for schedule(static)
        for (int t = 0; t < T; t++) {
            float *fr = frames + (size_t)t * n;
            irfft_frame(p, spec + (size_t)t * K, fr, w0, w1);
            for (int i = 0; i < n; i++) fr[i] *= p->win[i];
        }
        free(w0); free(w1);
#pragma omp for schedule(static)
        for (int j = 0; j < len; j++) {
            const int pos = j + half;             /* position in the untrimmed buffer */
            int t_hi = pos / hop; if (t_hi > T - 1) t_hi = T - 1;
            int t_lo = (pos - n) / hop + 1; if (pos - n < 0) t_lo = 0;
            float acc = 0.f;
            for (int t = t_lo; t <= t_hi; t++) acc += frames[(size_t)t * n + (pos - t * hop)];
            const float w = wss[pos];
            y[j] = (w > 1.17549435e-38f) ? acc / w : acc;
        }
    }
}

static void window_sumsquare(const plan_t *p, int hop, int T, float *wss) {
    const int n = p->n, total = n + hop * (T - 1);
    memset(wss, 0, sizeof(float) * total);
    for (int t = 0; t < T; t++)
        for (int i = 0; i < n; i++) wss[t * hop + i] += p->win[i] * p->win[i];
}

/* --------------------------------------------------------------- public API */
/* Griffin-Lim from a linear magnitude.  s_mag, turns: [K][T] row-major (the
 * reference's ndarray layout).  out: hop*(T-1) samples.  Returns 0. */
int oracle_gl_from_mag(const float *s_mag, const float *turns, int K, int T, int hop, int n_iter,
                       float momentum, int pad_constant, float *out) {
    const int n = 2 * (K - 1), len = hop * (T - 1);
    plan_t *p = plan_new(n);
    const size_t KT = (size_t)K * T;
    float *S = (float *)malloc(sizeof(float) * KT);        /* frame-major copy */
    cpx *ang = (cpx *)malloc(sizeof(cpx) * KT), *reb = (cpx *)calloc(KT, sizeof(cpx));
    cpx *prev = (cpx *)calloc(KT, sizeof(cpx)), *spec = (cpx *)malloc(sizeof(cpx) * KT);
    float *frames = (float *)malloc(sizeof(float) * (size_t)T * n);
    float *wss = (float *)malloc(sizeof(float) * (n + hop * (T - 1)));
    float *y = (float *)malloc(sizeof(float) * len);
    window_sumsquare(p, hop, T, wss);
    const float alpha = momentum / (1.0f + momentum), tiny = 1.17549435e-38f;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; t++)
        for (int k = 0; k < K; k++) {
            S[(size_t)t * K + k] = s_mag[(size_t)k * T + t];
            const double th = 2.0 * M_PI * (double)turns[(size_t)k * T + t];
            ang[(size_t)t * K + k].re = (float)cos(th); ang[(size_t)t * K + k].im = (float)sin(th);
        }
    for (int it = 0; it <= n_iter; it++) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < KT; i++) { spec[i].re = S[i] * ang[i].re; spec[i].im = S[i] * ang[i].im; }
        istft_fm(p, spec, hop, T, y, frames, wss);
        if (it == n_iter) break;
        { cpx *t = prev; prev = reb; reb = t; }
        stft_fm(p, y, len, hop, T, reb, pad_constant);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < KT; i++) {
            const float ar = reb[i].re - alpha * prev[i].re, ai = reb[i].im - alpha * prev[i].im;
            const float d = hypotf(ar, ai) + tiny;
            ang[i].re = ar / d; ang[i].im = ai / d;
        }
    }
    memcpy(out, y, sizeof(float) * len);
    free(S); free(ang); free(reb); free(prev); free(spec); free(frames); free(wss); free(y);
    plan_free(p);
    return 0;
}

/* S[K][T] = max(0, pinv[K][M] . delog(mel[M][T]))^power ; delog: 0 exp, 1 pow10, 2 none */
int oracle_lift_pinv(const float *pinv, const float *mel, int K, int M, int T, float power, int delog, float *s_mag) {
    float *e = (float *)malloc(sizeof(float) * (size_t)M * T);
    for (size_t i = 0; i < (size_t)M * T; i++)
        e[i] = delog == 0 ? expf(mel[i]) : (delog == 1 ? powf(10.f, mel[i]) : mel[i]);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; k++) {
        float *row = s_mag + (size_t)k * T;
        for (int t = 0; t < T; t++) row[t] = 0.f;
        for (int m = 0; m < M; m++) {
            const float c = pinv[(size_t)k * M + m];
            const float *er = e + (size_t)m * T;
            for (int t = 0; t < T; t++) row[t] += c * er[t];
        }
        for (int t = 0; t < T; t++) row[t] = powf(row[t] > 0.f ? row[t] : 0.f, power);
    }
    free(e);
    return 0;
}

/* GriffinLim::infer (src/lib.rs:141): mel [M][T] -> samples, peak-normalised when normalise==0 */
int oracle_infer(const float *pinv, const float *mel, const float *turns, int K, int M, int T, int hop,
                 float power, int n_iter, float momentum, int delog, int pad_constant, int normalise, float *out) {
    float *S = (float *)malloc(sizeof(float) * (size_t)K * T);
    oracle_lift_pinv(pinv, mel, K, M, T, power, delog, S);
    oracle_gl_from_mag(S, turns, K, T, hop, n_iter, momentum, pad_constant, out);
    free(S);
    if (normalise == 0) {
        const int len = hop * (T - 1);
        float mx = 0.f;
        for (int i = 0; i < len; i++) { const float a = fabsf(out[i]); if (a > mx) mx = a; }
        if (mx > 0.f) for (int i = 0; i < len; i++) out[i] /= mx;
    }
    return 0;
}

/* B utterances of equal shape, one OpenMP thread per utterance (the inner loops then run
 * serially: nested parallelism is off by default).  The better of this and the per-frame
 * parallel form is what bench.py reports as the CPU arm. */
int oracle_infer_batch(const float *pinv, const float *mels, const float *turns, int B, int K, int M, int T, int hop,
                       float power, int n_iter, float momentum, int delog, int pad_constant, int normalise, float *outs) {
    const size_t len = (size_t)hop * (T - 1);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; b++)
        oracle_infer(pinv, mels + (size_t)b * M * T, turns + (size_t)b * K * T, K, M, T, hop, power, n_iter, momentum,
                     delog, pad_constant, normalise, outs + b * len);
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* raw transforms, exported so tests can pin this file's FFT against numpy */
int oracle_stft(const float *y, int len, int n_fft, int hop, int pad_constant, float *spec_fm /* [T][K][2] */) {
    plan_t *p = plan_new(n_fft);
    const int T = 1 + len / hop;
    stft_fm(p, y, len, hop, T, (cpx *)spec_fm, pad_constant);
    plan_free(p);
    return T;
}
