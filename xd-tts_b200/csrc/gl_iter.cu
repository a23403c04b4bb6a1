// Griffin-Lim iteration kernel for sm_100a: one launch == one iteration of
//   rebuilt = stft(istft(S * angles)); angles = unit(rebuilt - alpha * tprev)
// (griffin_lim::GriffinLim::infer, external crate called at /root/reference src/lib.rs:141;
// see gl_core.cuh for the algorithm and the lane mapping).
//
// Launch shape: one warp per run of consecutive frames (table built on the host so that the
// whole batch is about one resident wave of warps), GL_WARPS warps per CTA that share only the
// read-only twiddle/window tables in shared memory.  No block-level barrier after the table
// load: warps are independent, the only cross-warp traffic is the two-party hand-over of the
// three hop blocks that straddle a run boundary (halo + arrival counter, second arriver sums).
//
// HBM traffic per frame and iteration (fp32): read S 4M, read R 8M, write R 8M, read y 4H,
// write y 4H  ~= 20K + 8H bytes; everything else stays in registers / smem.
#include <cuda_runtime.h>

#include "gl_core.cuh"
#include "gl_host.h"

namespace xdtts {

// warps per CTA and CTAs per SM by geometry: shared memory (tables + per-warp exchange and staging) is the limiter
// (tuning builds override the defaults with -DXDTTS_GL8_WARPS=... etc., tools/build_variants.py)
#ifndef XDTTS_GL8_WARPS
#define XDTTS_GL8_WARPS 4
#define XDTTS_GL8_CTAS 3
#define XDTTS_GL8_ALIAS 0
#endif
// n_fft 2048: one CTA of 8 warps per SM (one copy of the 21 KB tables instead of two: 451 -> 433 us per cfg5 launch)
#ifndef XDTTS_GL16_WARPS
#define XDTTS_GL16_WARPS 8
#define XDTTS_GL16_CTAS 1
#define XDTTS_GL16_ALIAS 1
#endif
#ifndef XDTTS_GL16_LATE
#define XDTTS_GL16_LATE 2
#endif
#ifndef XDTTS_GL8_LATE
#define XDTTS_GL8_LATE 0
#endif
template <int R3>
struct GlCfg {
    static constexpr int WARPS = (R3 == 16) ? XDTTS_GL16_WARPS : XDTTS_GL8_WARPS;
    static constexpr int CTAS = (R3 == 16) ? XDTTS_GL16_CTAS : XDTTS_GL8_CTAS;
    static constexpr bool ALIAS = (R3 == 16) ? XDTTS_GL16_ALIAS : XDTTS_GL8_ALIAS;   // exchange 1 and 2 share storage
    // when the next frame's newest hop block is requested: right after F1 (live through the whole frame), or before
    // the last inverse pass (live through a fifth of it: still several DRAM latencies, and out of the way of the
    // register peak in the middle passes)
    static constexpr int LATE_PREFETCH = (R3 == 16) ? XDTTS_GL16_LATE : XDTTS_GL8_LATE;   // 0: after F1, 2: after F3, 1: after F4
};

// ---- run-boundary hand-over.  The three hop blocks that straddle two runs get partial sums from both; the
// arrival counter of the boundary (flags[], +2 per launch, odd = one side is in) decides who finishes them.
// A run reaches its THIRD frame (head side done) long before its left neighbour reaches its LAST one, so the
// usual order is head first, tail second:
//   head  publishes its partial blocks (fence), bumps the counter and moves on; the counter value it got back is
//         looked at only when the run ends, so the atomic's round trip is off the warp's critical path;
//   tail  reads the counter (acquire); odd -> the head is in: it adds the head's partial sums to its register
//         accumulators and stores the finished blocks -- its own partial sums never go to memory, no fence, no
//         atomic with a result (a fire-and-forget red.add restores the counter's parity);
//         even -> the old two-party protocol: publish, fence, atomic; whoever comes second adds left + right.
// All three paths add (left + right) in that order: bit-identical results.
__device__ __forceinline__ void fence_release_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// head side: returns the counter value before this arrival (valid in every lane)
__device__ __forceinline__ unsigned arrive_head(int lane, const GlParams& p, int boundary) {
    fence_release_gpu();   // publish this lane's partial sums
    __syncwarp();
    unsigned old = 0;
    if (lane == 0) old = atomicAdd(p.flags + boundary, 1u);
    return old;            // lane 0's value is broadcast when it is consumed (end of the run)
}

template <int R3, bool TRACK_MAX>
__device__ __forceinline__ void arrive(Lane<R3>& L, int lane, const GlParams& p, int boundary) {
    fence_release_gpu();   // publish this lane's partial sums
    __syncwarp();
    unsigned old = 0;
    if (lane == 0) old = atomicAdd(p.flags + boundary, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old & 1u) {    // the neighbour was here first: its partial sums are visible after the fence
        fence_release_gpu();
        combine_boundary<R3, TRACK_MAX>(L, lane, p, boundary);
    }
}

// shared memory of one CTA, float2 units: [tables][per warp: exchange 1, exchange 2, staged state record (R | S | S_nyq)]
// followed by the mbarriers (one for the table copy, one per warp for its state staging)
template <int R3>
struct GlSmem {
    typedef Geo<R3> G;
    static constexpr int WARPS = GlCfg<R3>::WARPS;
    static constexpr int EX = GlCfg<R3>::ALIAS ? (G::EX1 > G::EX2 ? G::EX1 : G::EX2) : G::EXW;
    static constexpr int WARP = EX + G::REC_F / 2;                  // float2 per warp: exchange scratch + the staged state record
    static constexpr int BAR_OFF = G::TAB + WARPS * WARP;           // float2 units (8 bytes each)
    static constexpr size_t BYTES = sizeof(float2) * (size_t)(BAR_OFF + 1 + WARPS);
};

// per-warp view of the CTA's shared memory
template <int R3>
struct WarpSmem {
    float2 *tab, *ex1, *ex2;
    float* stg;   // staged state record of one frame: [R | S | S_nyq]
    unsigned long long *bar_tab, *bar;
    __device__ __forceinline__ WarpSmem(float2* smem, int warp) {
        typedef Geo<R3> G;
        tab = smem;
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + GlSmem<R3>::BAR_OFF);
        ex1 = smem + G::TAB + warp * GlSmem<R3>::WARP;
        ex2 = GlCfg<R3>::ALIAS ? ex1 : ex1 + G::EX1;
        stg = reinterpret_cast<float*>(ex1 + GlSmem<R3>::EX);
        bar_tab = &bars[0];
        bar = &bars[1 + warp];
    }
};

// CTA prologue: barriers + the bulk copy of the constant tables (awaited by each warp before its first use)
template <int R3>
__device__ __forceinline__ void cta_prologue(float2* smem, const GlParams& p) {
    typedef Geo<R3> G;
    constexpr int GL_WARPS = GlCfg<R3>::WARPS;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + GlSmem<R3>::BAR_OFF);
    if (threadIdx.x == 0) {
        for (int i = 0; i <= GL_WARPS; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], (unsigned)(G::TAB * sizeof(float2)));
        bulk_g2s(smem, p.tables, (unsigned)(G::TAB * sizeof(float2)), &bars[0]);
    }
}

// One Griffin-Lim iteration over the frames of one run.  `phase` is the parity of the warp's staging
// barrier (it flips once per frame and carries over between iterations in the persistent kernel).
// COHERENT: the previous waveform was written by other SMs during this launch -> read it through L2.
template <int R3, int MODE, bool STORE_R, bool TRACK_MAX, bool COHERENT>
__device__ __forceinline__ void gl_run_frames(Lane<R3>& L, int lane, const GlParams& p, int run_idx, const GlRun& r, int T,
                                              long foff, const WarpSmem<R3>& sm, unsigned& phase) {
    typedef Geo<R3> G;
    constexpr bool ALIAS = GlCfg<R3>::ALIAS;
    const long yoff = foff * G::H;
    float2 *tab = sm.tab, *ex1 = sm.ex1, *ex2 = sm.ex2;
    stage_issue<R3, MODE>(lane, p, foff + r.ta, sm.stg, sm.bar);
    bool pref = false;
    unsigned head_old = 0;        // counter value the head-side arrival saw (lane 0), examined after the loop
    const bool has_left = r.ta > 0, has_right = r.tb < T;
    for (int t = r.ta; t < r.tb; t++) {
        const long frame = foff + t;
        if (MODE != GL_MODE_INIT) {
            const bool fetch_next = (t + 1 < r.tb) && (t + 2 <= T - 2);   // newest hop block of frame t+1 lies inside the signal
            if constexpr (G::ROT) {
                phase_f1_rot<R3, COHERENT>(L, lane, p.y_in + yoff, T, t, p.pad_mode, t == r.ta, pref, fetch_next, (t - r.ta) & 3, tab, ex1);
                pref = fetch_next;
                __syncwarp();
            } else {
                phase_f1<R3, COHERENT>(L, lane, p.y_in + yoff, T, t, p.pad_mode, t == r.ta, pref, tab, ex1);
                pref = fetch_next;
                __syncwarp();
                if (!GlCfg<R3>::LATE_PREFETCH && fetch_next) prefetch_next_block<R3, COHERENT>(L, lane, p.y_in + yoff, T, t, p.pad_mode);
            }
            phase_f2_load<R3>(L, lane, ex1);
            if (ALIAS) __syncwarp();
            phase_f2_store<R3>(L, lane, tab, ex2);
            __syncwarp();
        }
        stage_wait(sm.bar, phase);
        phase ^= 1u;
        phase_f3<R3, MODE, STORE_R>(L, lane, p, r.utt, T, t, frame, tab, ex2, sm.stg);
        __syncwarp();   // every lane is done with the staged state (its values fed the stores above)
        if (t + 1 < r.tb) stage_issue<R3, MODE>(lane, p, frame + 1, sm.stg, sm.bar);
        if (!G::ROT && MODE != GL_MODE_INIT && GlCfg<R3>::LATE_PREFETCH == 2 && pref) prefetch_next_block<R3, COHERENT>(L, lane, p.y_in + yoff, T, t, p.pad_mode);
        phase_f4_load<R3>(L, lane, tab, ex2);
        if (ALIAS) __syncwarp();
        phase_f4_store<R3>(L, lane, ex1);
        __syncwarp();
        if (!G::ROT && MODE != GL_MODE_INIT && GlCfg<R3>::LATE_PREFETCH == 1 && pref) prefetch_next_block<R3, COHERENT>(L, lane, p.y_in + yoff, T, t, p.pad_mode);
        phase_f5<R3>(L, lane, tab, ex1);
        float2 out[2 * G::NB];
        ola_shift<R3>(L, out);
        if (emit_block<R3, TRACK_MAX>(L, lane, p, run_idx, r, yoff, t, out)) head_old = arrive_head(lane, p, run_idx - 1);
        // the next frame's first shared-memory writes (F1 -> ex1) go to the addresses this same lane read last in F5,
        // so no barrier is needed here -- except in INIT mode with aliased exchange buffers, where the next frame
        // starts with F3's writes to ex2 (= ex1) in a different index mapping than F5's reads (compute-sanitizer
        // racecheck flagged exactly this pair)
        if (ALIAS && MODE == GL_MODE_INIT) __syncwarp();
    }
    if (has_right) {
        const bool ready = (ld_acquire_gpu(p.flags + run_idx) & 1u) != 0;   // same address in every lane: one request
        if (emit_tail<R3, TRACK_MAX>(L, lane, p, run_idx, r, yoff, T, ready)) {
            arrive<R3, TRACK_MAX>(L, lane, p, run_idx);
        } else if (lane == 0) {
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + run_idx) : "memory");
        }
    } else {
        emit_tail<R3, TRACK_MAX>(L, lane, p, run_idx, r, yoff, T);
    }
    if (has_left) {   // the left neighbour finished its run before this one reached its third frame (rare)
        head_old = __shfl_sync(0xffffffffu, head_old, 0);
        if (head_old & 1u) {
            fence_release_gpu();
            combine_boundary<R3, TRACK_MAX>(L, lane, p, run_idx - 1);
        }
    }

    if (TRACK_MAX) {
        float m = L.amax;
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(p.amax + r.utt, __float_as_uint(m));
    }
}

template <int R3, int MODE, bool STORE_R, bool TRACK_MAX>
__global__ void __launch_bounds__(GlCfg<R3>::WARPS * 32, GlCfg<R3>::CTAS) gl_iter_kernel(const GlParams p) {
    extern __shared__ __align__(16) float2 smem[];
    cta_prologue<R3>(smem, p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int run_idx = blockIdx.x * GlCfg<R3>::WARPS + warp;
    if (run_idx >= p.n_runs) return;
    const WarpSmem<R3> sm(smem, warp);
    const GlRun r = p.runs[run_idx];
    const int T = p.utt_T[r.utt];
    const long foff = p.utt_foff[r.utt];
    Lane<R3> L;
    lane_reset<R3>(L);
    stage_wait(sm.bar_tab, 0);
    lane_load_constants<R3>(L, lane, sm.tab);
    unsigned phase = 0;
    gl_run_frames<R3, MODE, STORE_R, TRACK_MAX, false>(L, lane, p, run_idx, r, T, foff, sm, phase);
}

// ------------------------------------------------------------------ persistent kernel
// The whole vocode -- initial inverse transform + n_iter iterations -- in ONE cooperative launch:
// every run keeps its warp, and a warp starts iteration i as soon as its two neighbouring runs (the
// only ones whose hop blocks it reads) have published iteration i-1, instead of waiting for the whole
// grid at a launch boundary.  The per-launch tail (SMs idling until the slowest warp is done) and the
// per-launch prologue disappear; iterations of different runs overlap.  Neighbours never drift more than
// one iteration apart, so two waveform buffers still suffice.  Needs every run resident at once: the host
// uses it only when the run table fits the device's resident warps (cooperative launch checks it).
__device__ __forceinline__ void wait_done(const unsigned* flag, unsigned want) {
    unsigned v;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v < want) __nanosleep(64);
    } while (v < want);
}

template <int R3>
__global__ void __launch_bounds__(GlCfg<R3>::WARPS * 32, GlCfg<R3>::CTAS) gl_persist_kernel(const GlParams p0) {
    extern __shared__ __align__(16) float2 smem[];
    cta_prologue<R3>(smem, p0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int run_idx = blockIdx.x * GlCfg<R3>::WARPS + warp;
    if (run_idx >= p0.n_runs) return;
    const WarpSmem<R3> sm(smem, warp);
    const GlRun r = p0.runs[run_idx];
    const int T = p0.utt_T[r.utt];
    const long foff = p0.utt_foff[r.utt];
    const bool has_left = r.ta > 0, has_right = r.tb < T;   // neighbouring runs of the same utterance
    Lane<R3> L;
    stage_wait(sm.bar_tab, 0);
    lane_load_constants<R3>(L, lane, sm.tab);
    unsigned phase = 0;
    GlParams p = p0;
    for (int it = 0; it <= p0.n_iter; it++) {
        p.y_in = p0.ybuf[(it + 1) & 1];
        p.y_out = p0.ybuf[it & 1];
        if (it > 0) {   // neighbours must have finished iteration it-1 (they also finished reading what this one overwrites)
            if (lane == 0) {
                if (has_left) wait_done(p0.done + run_idx - 1, (unsigned)it);
                if (has_right) wait_done(p0.done + run_idx + 1, (unsigned)it);
            }
            __syncwarp();
        }
        lane_reset<R3>(L);
        const bool last = it == p0.n_iter;
        if (it == 0) {
            if (last) gl_run_frames<R3, GL_MODE_INIT, false, true, true>(L, lane, p, run_idx, r, T, foff, sm, phase);
            else gl_run_frames<R3, GL_MODE_INIT, false, false, true>(L, lane, p, run_idx, r, T, foff, sm, phase);
        } else if (it == 1) {
            if (last) gl_run_frames<R3, GL_MODE_FIRST, false, true, true>(L, lane, p, run_idx, r, T, foff, sm, phase);
            else gl_run_frames<R3, GL_MODE_FIRST, true, false, true>(L, lane, p, run_idx, r, T, foff, sm, phase);
        } else {
            if (last) gl_run_frames<R3, GL_MODE_MID, false, true, true>(L, lane, p, run_idx, r, T, foff, sm, phase);
            else gl_run_frames<R3, GL_MODE_MID, true, false, true>(L, lane, p, run_idx, r, T, foff, sm, phase);
        }
        // publish: this run's waveform blocks and rebuilt spectrum of iteration `it` are complete.  The
        // spectrum rows are re-read by this warp's bulk copies (async proxy) in the next iteration.
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p0.done + run_idx), "r"((unsigned)(it + 1)) : "memory");
    }
}

template <int R3>
static size_t gl_smem_bytes() {
    return GlSmem<R3>::BYTES;
}

template <int R3>
static cudaError_t persist_prepare() {
    return cudaFuncSetAttribute(gl_persist_kernel<R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gl_smem_bytes<R3>());
}

template <int R3>
static cudaError_t persist_launch(const GlParams& p, int sm_count, cudaStream_t s, bool* fits) {
    constexpr int W = GlCfg<R3>::WARPS;
    const int grid = (p.n_runs + W - 1) / W;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gl_persist_kernel<R3>, W * 32, gl_smem_bytes<R3>());
    if (e != cudaSuccess) return e;
    *fits = grid <= per_sm * sm_count;
    if (!*fits) return cudaSuccess;
    GlParams q = p;
    void* args[1] = {&q};
    return cudaLaunchCooperativeKernel((const void*)gl_persist_kernel<R3>, dim3(grid), dim3(W * 32), args, gl_smem_bytes<R3>(), s);
}

// the whole vocode in one cooperative launch; *fits == false (and nothing launched) when the run table
// exceeds the resident warps of the device
cudaError_t gl_launch_persistent(int n_fft, const GlParams& p, int sm_count, cudaStream_t s, bool* fits) {
    switch (n_fft) {
        case 512: return persist_launch<4>(p, sm_count, s, fits);
        case 1024: return persist_launch<8>(p, sm_count, s, fits);
        case 2048: return persist_launch<16>(p, sm_count, s, fits);
    }
    return cudaErrorInvalidValue;
}

template <int R3, int MODE, bool STORE_R, bool TRACK_MAX>
static cudaError_t prepare_one() {
    return cudaFuncSetAttribute(gl_iter_kernel<R3, MODE, STORE_R, TRACK_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)gl_smem_bytes<R3>());
}

template <int R3>
static cudaError_t prepare_r3() {
    cudaError_t e = prepare_one<R3, GL_MODE_INIT, false, true>();
    if (e == cudaSuccess) e = prepare_one<R3, GL_MODE_INIT, false, false>();
    if (e == cudaSuccess) e = prepare_one<R3, GL_MODE_FIRST, false, true>();
    if (e == cudaSuccess) e = prepare_one<R3, GL_MODE_FIRST, true, false>();
    if (e == cudaSuccess) e = prepare_one<R3, GL_MODE_MID, false, true>();
    if (e == cudaSuccess) e = prepare_one<R3, GL_MODE_MID, true, false>();
    if (e == cudaSuccess) e = persist_prepare<R3>();
    return e;
}

// opt the kernels of this geometry into their dynamic shared memory size on the current device
// (called once per handle, outside any stream capture)
cudaError_t gl_prepare(int n_fft) {
    switch (n_fft) {
        case 512: return prepare_r3<4>();
        case 1024: return prepare_r3<8>();
        case 2048: return prepare_r3<16>();
    }
    return cudaErrorInvalidValue;
}

template <int R3, int MODE, bool STORE_R, bool TRACK_MAX>
static cudaError_t launch_one(const GlParams& p, cudaStream_t s) {
    auto k = gl_iter_kernel<R3, MODE, STORE_R, TRACK_MAX>;
    constexpr int W = GlCfg<R3>::WARPS;
    const int grid = (p.n_runs + W - 1) / W;
    k<<<grid, W * 32, gl_smem_bytes<R3>(), s>>>(p);
    return cudaGetLastError();
}

template <int R3>
static cudaError_t launch_r3(int mode, bool last, const GlParams& p, cudaStream_t s) {
    // the last launch of a sequence needs no R store and tracks max |y| for the peak normalisation
    if (mode == GL_MODE_INIT) {
        return last ? launch_one<R3, GL_MODE_INIT, false, true>(p, s) : launch_one<R3, GL_MODE_INIT, false, false>(p, s);
    } else if (mode == GL_MODE_FIRST) {
        return last ? launch_one<R3, GL_MODE_FIRST, false, true>(p, s) : launch_one<R3, GL_MODE_FIRST, true, false>(p, s);
    }
    return last ? launch_one<R3, GL_MODE_MID, false, true>(p, s) : launch_one<R3, GL_MODE_MID, true, false>(p, s);
}

cudaError_t gl_launch_iteration(int n_fft, int mode, bool last, const GlParams& p, cudaStream_t s) {
    switch (n_fft) {
        case 512: return launch_r3<4>(mode, last, p, s);
        case 1024: return launch_r3<8>(mode, last, p, s);
        case 2048: return launch_r3<16>(mode, last, p, s);
    }
    return cudaErrorInvalidValue;
}

int gl_warps_per_cta(int n_fft) { return n_fft == 2048 ? GlCfg<16>::WARPS : GlCfg<8>::WARPS; }

int gl_ctas_per_sm(int n_fft) { return n_fft == 2048 ? GlCfg<16>::CTAS : GlCfg<8>::CTAS; }

int gl_resident_warps_per_sm(int n_fft) {
    return n_fft == 2048 ? GlCfg<16>::WARPS * GlCfg<16>::CTAS : GlCfg<8>::WARPS * GlCfg<8>::CTAS;
}

}  // namespace xdtts
