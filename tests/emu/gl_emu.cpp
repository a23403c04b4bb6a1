// CPU emulation of the Griffin-Lim iteration kernel's lane program -- TEST INFRASTRUCTURE.
//
// Runs xd-tts_b200/csrc/gl_core.cuh (the very functions the CUDA kernel inlines) with the 32
// lanes of a warp looped on the host and a "warp sync" between phases, so the index algebra
// (FFT factorisation, exchange layouts, pairing, run boundaries, edge frames) can be checked
// against the oracle in the GPU-less build container.  Never loaded by the product.
//
// build: g++ -O2 -shared -fPIC -I/usr/local/cuda/include -Ixd-tts_b200/csrc tests/emu/gl_emu.cpp
#include <cmath>
#include <cstring>
#include <vector>

#include "gl_core.cuh"
#include "gl_tables.h"

using namespace xdtts;

template <int R3, int MODE, bool STORE_R, bool TRACK_MAX>
static void emu_launch(GlParams p, bool reverse_order) {
    typedef Geo<R3> G;
    // one buffer for both exchanges: the emulator always runs the aliased layout (a superset of the
    // hazards of the two-buffer layout; phases are lane-sequential, i.e. warp-synchronous)
    std::vector<float2> ex1(G::EX1 > G::EX2 ? G::EX1 : G::EX2);
    std::vector<float2>& ex2 = ex1;
    std::vector<float> stg(G::REC_F);   // staged state record [R | S | S_nyq]
    std::vector<Lane<R3>> lanes(32);
    for (int rr = 0; rr < p.n_runs; rr++) {
        const int run_idx = reverse_order ? p.n_runs - 1 - rr : rr;
        const GlRun r = p.runs[run_idx];
        const int T = p.utt_T[r.utt];
        const long foff = p.utt_foff[r.utt];
        const long yoff = foff * G::H;
        for (int l = 0; l < 32; l++) {
            lane_reset<R3>(lanes[l]);
            lane_load_constants<R3>(lanes[l], l, p.tables);
        }
        auto arrive = [&](int boundary) {
            const unsigned old = p.flags[boundary]++;
            if (old & 1u)
                for (int l = 0; l < 32; l++) combine_boundary<R3, TRACK_MAX>(lanes[l], l, p, boundary);
        };
        for (int l = 0; l < 32; l++) stage_issue<R3, MODE>(l, p, foff + r.ta, stg.data(), nullptr);
        bool pref = false;
        for (int t = r.ta; t < r.tb; t++) {
            const long frame = foff + t;
            if (MODE != GL_MODE_INIT) {
                const bool fetch_next = (t + 1 < r.tb) && (t + 2 <= T - 2);
                if (G::ROT) {
                    for (int l = 0; l < 32; l++)
                        phase_f1_rot<R3>(lanes[l], l, p.y_in + yoff, T, t, p.pad_mode, t == r.ta, pref, fetch_next, (t - r.ta) & 3,
                                         p.tables, ex1.data());
                    pref = fetch_next;
                } else {
                    for (int l = 0; l < 32; l++)
                        phase_f1<R3>(lanes[l], l, p.y_in + yoff, T, t, p.pad_mode, t == r.ta, pref, p.tables, ex1.data());
                    pref = fetch_next;
                    if (fetch_next)
                        for (int l = 0; l < 32; l++) prefetch_next_block<R3>(lanes[l], l, p.y_in + yoff, T, t, p.pad_mode);
                }
                for (int l = 0; l < 32; l++) phase_f2_load<R3>(lanes[l], l, ex1.data());
                for (int l = 0; l < 32; l++) phase_f2_store<R3>(lanes[l], l, p.tables, ex2.data());
            }
            for (int l = 0; l < 32; l++)
                phase_f3<R3, MODE, STORE_R>(lanes[l], l, p, r.utt, T, t, frame, p.tables, ex2.data(), stg.data());
            if (t + 1 < r.tb)
                for (int l = 0; l < 32; l++) stage_issue<R3, MODE>(l, p, frame + 1, stg.data(), nullptr);
            for (int l = 0; l < 32; l++) phase_f4_load<R3>(lanes[l], l, p.tables, ex2.data());
            for (int l = 0; l < 32; l++) phase_f4_store<R3>(lanes[l], l, ex1.data());
            bool sig = false;
            for (int l = 0; l < 32; l++) {
                phase_f5<R3>(lanes[l], l, p.tables, ex1.data());
                float2 out[2 * G::NB];
                ola_shift<R3>(lanes[l], out);
                sig = emit_block<R3, TRACK_MAX>(lanes[l], l, p, run_idx, r, yoff, t, out);
            }
            if (sig) arrive(run_idx - 1);
        }
        // the kernel's tail-side fast path (gl_iter.cu): the right neighbour is already in (counter odd) -> finish
        // the boundary blocks from the register accumulators and only restore the counter's parity
        const bool ready = r.tb < T && (p.flags[run_idx] & 1u);
        bool sig = false;
        for (int l = 0; l < 32; l++) sig = emit_tail<R3, TRACK_MAX>(lanes[l], l, p, run_idx, r, yoff, T, ready);
        if (sig) arrive(run_idx);
        else if (ready) p.flags[run_idx]++;
        if (TRACK_MAX) {
            float m = 0.f;
            for (int l = 0; l < 32; l++) m = fmaxf(m, lanes[l].amax);
            unsigned bits;
            memcpy(&bits, &m, 4);
            if (bits > p.amax[r.utt]) p.amax[r.utt] = bits;
        }
    }
}

template <int R3>
static int emu_run(const float* s_mag, const float* turns, int T, int n_iter, float momentum, int run_frames, int pad_mode,
                   unsigned long long seed, float* out, float* r_out, float* peak) {
    typedef Geo<R3> G;
    const int K = G::M + 1;
    std::vector<GlRun> runs;
    std::vector<int> foff;
    build_runs(&T, 1, run_frames, &runs, &foff);
    std::vector<float2> tab = build_tables<R3>();
    std::vector<float> edge = build_edge_scale(G::N);
    // per-frame state records [R: 2M floats | S: M floats | S_nyq | 3 pad] and the initial phase [frames][M + 1]
    std::vector<float> state((size_t)T * G::REC_F, 0.f), tu((size_t)T * K);
    for (int t = 0; t < T; t++)
        for (int k = 0; k < K; k++) {
            state[(size_t)t * G::REC_F + 2 * G::M + k] = s_mag[(size_t)k * T + t];
            if (turns) tu[(size_t)t * K + k] = turns[(size_t)k * T + t];
        }
    std::vector<float> ya((size_t)T * G::H, 0.f), yb((size_t)T * G::H, 0.f);
    std::vector<float> halo((size_t)runs.size() * 6 * G::H, 0.f);
    std::vector<unsigned> flags(runs.size(), 0u);
    unsigned amax = 0;
    GlParams p;
    memset(&p, 0, sizeof(p));
    p.n_runs = (int)runs.size();
    p.runs = runs.data();
    p.utt_T = &T;
    p.utt_foff = foff.data();
    p.state = state.data();
    p.halo = halo.data();
    p.flags = flags.data();
    p.amax = &amax;
    p.edge_scale = edge.data();
    p.tables = tab.data();
    p.turns = turns ? tu.data() : nullptr;
    const int seed_id = 0;
    p.seed = &seed;
    p.utt_seed_id = &seed_id;
    p.alpha = momentum / (1.0f + momentum);
    p.inv_n = 1.0f / (float)G::N;
    p.pad_mode = pad_mode;
    float* yin = ya.data();
    float* yout = yb.data();
    p.y_in = yin;
    p.y_out = yout;
    if (n_iter == 0) emu_launch<R3, GL_MODE_INIT, false, true>(p, false);
    else emu_launch<R3, GL_MODE_INIT, false, false>(p, false);
    for (int it = 1; it <= n_iter; it++) {
        std::swap(yin, yout);
        p.y_in = yin;
        p.y_out = yout;
        const bool last = (it == n_iter);
        const bool rev = (it & 1);   // alternate the processing order: exercises both arrival orders
        if (it == 1) {
            if (last) emu_launch<R3, GL_MODE_FIRST, false, true>(p, rev);
            else emu_launch<R3, GL_MODE_FIRST, true, false>(p, rev);
        } else {
            if (last) emu_launch<R3, GL_MODE_MID, false, true>(p, rev);
            else emu_launch<R3, GL_MODE_MID, true, false>(p, rev);
        }
    }
    memcpy(out, yout, sizeof(float) * (size_t)G::H * (T - 1));
    if (r_out)
        for (int t = 0; t < T; t++) memcpy(r_out + (size_t)t * 2 * G::M, &state[(size_t)t * G::REC_F], sizeof(float) * 2 * G::M);
    if (peak) memcpy(peak, &amax, 4);
    return (int)runs.size();
}

extern "C" int emu_gl_from_mag(const float* s_mag, const float* turns, int K, int T, int n_iter, float momentum,
                               int run_frames, int pad_mode, unsigned long long seed, float* out, float* r_out,
                               float* peak) {
    if (T < 4) return -1;
    switch (K) {
        case 257: return emu_run<4>(s_mag, turns, T, n_iter, momentum, run_frames, pad_mode, seed, out, r_out, peak);
        case 513: return emu_run<8>(s_mag, turns, T, n_iter, momentum, run_frames, pad_mode, seed, out, r_out, peak);
        case 1025: return emu_run<16>(s_mag, turns, T, n_iter, momentum, run_frames, pad_mode, seed, out, r_out, peak);
    }
    return -2;
}
