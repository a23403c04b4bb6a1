// Host-side construction of the constant tables and the run table consumed by the
// Griffin-Lim iteration kernel (gl_core.cuh).  Pure C++ (double precision), shared by the
// CUDA library and by the CPU lane-program emulator under tests/emu/.
#pragma once
#include <cmath>
#include <vector>

#include "gl_core.cuh"

namespace xdtts {

// periodic Hann, w[n] = 0.5 - 0.5 cos(2 pi n / N): librosa get_window('hann', N, fftbins=True)
// (the crate behind /root/reference src/tacotron2/mod.rs:456 ports librosa)
inline double hann_periodic(int n, int N) { return 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)N); }

template <int R3>
std::vector<float2> build_tables() {
    typedef Geo<R3> G;
    std::vector<float2> t(G::TAB);
    // paired rows: entry `pair` of a block holds rows (2 pair, 2 pair + 1) of one lane in 16 consecutive bytes (gl_core.cuh tab_pair)
    auto put_pair = [&](int base, int row, int lane, float2 v) { t[base + 64 * (row >> 1) + 2 * lane + (row & 1)] = v; };
    for (int lane = 0; lane < 32; lane++) {
        for (int i = 0; i < G::NB; i++)
            for (int k1 = 1; k1 < 8; k1++) {
                const double a = -2.0 * M_PI * (double)((lane + 32 * i) * k1) / (double)G::M;
                const float2 v = mk2((float)std::cos(a), (float)std::sin(a));
                const int base = G::TW1_OFF + i * 7 * 32;
                if (k1 < 7) put_pair(base, k1 - 1, lane, v);       // (1,2) (3,4) (5,6)
                else t[base + 6 * 32 + lane] = v;                  // k1 = 7 alone
            }
        for (int k2 = 1; k2 < 8; k2++) {
            const double a = -2.0 * M_PI * (double)((lane & (R3 - 1)) * k2) / (double)(8 * R3);
            t[G::TW2_OFF + (k2 - 1) * 32 + lane] = mk2((float)std::cos(a), (float)std::sin(a));
        }
        for (int j = 0; j < R3; j++) {
            const int k = kslot<R3>(lane, j);
            const double a = 2.0 * M_PI * (double)k / (double)G::N;   // W_N^k = cos a - i sin a
            put_pair(G::RTW_OFF, j, lane, mk2((float)(-0.5 * std::sin(a)), (float)(-0.5 * std::cos(a))));
        }
        for (int i = 0; i < G::NB; i++)
            for (int n1 = 0; n1 < 4; n1++) {   // first half of the window only: w[s + N/2] = 1 - w[s]
                const int s = 16 * R3 * n1 + 2 * (lane + 32 * i);
                put_pair(G::WIN_OFF + i * 4 * 32, n1, lane, mk2((float)hann_periodic(s, G::N), (float)hann_periodic(s + 1, G::N)));
            }
    }
    return t;
}

// 1 / sum_t w^2 for the first (frames 0,1,2) and last (frames T-3,T-2,T-1) hop block of the
// trimmed signal; every other block sees four frames and sums to 1.5 exactly
inline std::vector<float> build_edge_scale(int N) {
    const int H = N / 4;
    std::vector<float> e(2 * H);
    for (int i = 0; i < H; i++) {
        float w0 = (float)hann_periodic(i, N), w1 = (float)hann_periodic(H + i, N);
        float w2 = (float)hann_periodic(2 * H + i, N), w3 = (float)hann_periodic(3 * H + i, N);
        // fp32 accumulation in frame order, as the reference-style window_sumsquare does
        float first = w2 * w2; first += w1 * w1; first += w0 * w0;
        float last = w3 * w3; last += w2 * w2; last += w1 * w1;
        e[i] = 1.0f / first;
        e[H + i] = 1.0f / last;
    }
    return e;
}

// split utterance u into counts[u] equal runs (each >= 4 frames so that a hop block is shared by at
// most two runs); runs of one utterance are consecutive in the table
inline void build_runs_counts(const int* T, int n_utt, const int* counts, std::vector<GlRun>* runs, std::vector<int>* foff) {
    runs->clear();
    foff->clear();
    int f = 0;
    for (int u = 0; u < n_utt; u++) {
        foff->push_back(f);
        f += T[u];
        int n = counts[u] < 1 ? 1 : counts[u];
        while (n > 1 && T[u] / n < 4) n--;
        int t = 0;
        for (int r = 0; r < n; r++) {
            const int len = T[u] / n + (r < T[u] % n ? 1 : 0);
            GlRun g;
            g.utt = u; g.ta = t; g.tb = t + len; g.pad = 0;
            runs->push_back(g);
            t += len;
        }
    }
}

// split every utterance into runs of about `run_frames` frames (each >= 4 so that a hop block is
// shared by at most two runs); runs of one utterance are consecutive in the table
inline void build_runs(const int* T, int n_utt, int run_frames, std::vector<GlRun>* runs, std::vector<int>* foff) {
    runs->clear();
    foff->clear();
    int f = 0;
    if (run_frames < 4) run_frames = 4;
    for (int u = 0; u < n_utt; u++) {
        foff->push_back(f);
        f += T[u];
        int n = (T[u] + run_frames - 1) / run_frames;
        if (n < 1) n = 1;
        // equalise: n runs of floor/ceil(T/n) frames
        while (n > 1 && T[u] / n < 4) n--;
        int t = 0;
        for (int r = 0; r < n; r++) {
            const int len = T[u] / n + (r < T[u] % n ? 1 : 0);
            GlRun g;
            g.utt = u; g.ta = t; g.tb = t + len; g.pad = 0;
            runs->push_back(g);
            t += len;
        }
    }
}

}  // namespace xdtts
