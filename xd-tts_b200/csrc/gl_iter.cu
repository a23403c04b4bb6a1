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

constexpr int GL_WARPS = 4;

template <int R3, bool TRACK_MAX>
__device__ __forceinline__ void arrive(Lane<R3>& L, int lane, const GlParams& p, int boundary) {
    __threadfence();   // publish this lane's partial sums
    __syncwarp();
    unsigned old = 0;
    if (lane == 0) old = atomicAdd(p.flags + boundary, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old & 1u) {    // the neighbour was here first: its partial sums are visible after the fence
        __threadfence();
        combine_boundary<R3, TRACK_MAX>(L, lane, p, boundary);
    }
}

template <int R3, int MODE, bool STORE_R, bool TRACK_MAX>
__global__ void __launch_bounds__(GL_WARPS * 32, (R3 == 16) ? 2 : 3) gl_iter_kernel(const GlParams p) {
    typedef Geo<R3> G;
    extern __shared__ __align__(16) float2 smem[];
    float2* tab = smem;
    for (int i = threadIdx.x; i < G::TAB; i += GL_WARPS * 32) tab[i] = p.tables[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int run_idx = blockIdx.x * GL_WARPS + warp;
    if (run_idx >= p.n_runs) return;
    float2* ex1 = smem + G::TAB + warp * G::EXW;
    float2* ex2 = ex1 + G::EX1;

    const GlRun r = p.runs[run_idx];
    const int T = p.utt_T[r.utt];
    const long foff = p.utt_foff[r.utt];
    const long yoff = foff * G::H;

    Lane<R3> L;
    lane_reset<R3>(L);

    for (int t = r.ta; t < r.tb; t++) {
        const long frame = foff + t;
        phase_f0<R3, MODE>(L, lane, p, frame);
        if (MODE != GL_MODE_INIT) {
            const bool have_pref = (t > r.ta) && !frame_is_edge(t, T);
            const bool fetch_next = (t + 1 < r.tb) && !frame_is_edge(t + 1, T);
            phase_f1<R3>(L, lane, p.y_in + yoff, T, t, p.pad_mode, have_pref, fetch_next, tab, ex1);
            __syncwarp();
            phase_f2<R3>(L, lane, tab, ex1, ex2);
            __syncwarp();
        }
        phase_f3<R3, MODE, STORE_R>(L, lane, p, r.utt, T, t, frame, tab, ex2);
        __syncwarp();
        phase_f4<R3>(L, lane, tab, ex2, ex1);
        __syncwarp();
        phase_f5<R3>(L, lane, tab, ex1);
        float2 out[2 * G::NB];
        ola_shift<R3>(L, out);
        if (emit_block<R3, TRACK_MAX>(L, lane, p, run_idx, r, yoff, t, out)) arrive<R3, TRACK_MAX>(L, lane, p, run_idx - 1);
        // the next frame's first shared-memory writes (F1 -> ex1, or F3 -> ex2 in INIT mode) go to the
        // addresses this same lane read last, so no barrier is needed here
    }
    if (emit_tail<R3, TRACK_MAX>(L, lane, p, run_idx, r, yoff, T)) arrive<R3, TRACK_MAX>(L, lane, p, run_idx);

    if (TRACK_MAX) {
        float m = L.amax;
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(p.amax + r.utt, __float_as_uint(m));
    }
}

template <int R3>
static size_t gl_smem_bytes() {
    return sizeof(float2) * (size_t)(Geo<R3>::TAB + GL_WARPS * Geo<R3>::EXW);
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
    const int grid = (p.n_runs + GL_WARPS - 1) / GL_WARPS;
    k<<<grid, GL_WARPS * 32, gl_smem_bytes<R3>(), s>>>(p);
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

int gl_warps_per_cta() { return GL_WARPS; }

int gl_resident_warps_per_sm(int n_fft) { return GL_WARPS * (n_fft == 2048 ? 2 : 3); }

}  // namespace xdtts
