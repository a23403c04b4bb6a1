// Internal host-side declarations shared by the CUDA translation units of libxdtts_b200.
#pragma once
#include <cuda_runtime.h>

#include "gl_core.cuh"

namespace xdtts {

// one Griffin-Lim iteration (gl_iter.cu); mode = GL_MODE_*, last = final launch of the sequence
cudaError_t gl_launch_iteration(int n_fft, int mode, bool last, const GlParams& p, cudaStream_t s);
cudaError_t gl_launch_persistent(int n_fft, const GlParams& p, int sm_count, cudaStream_t s, bool* fits);
cudaError_t gl_prepare(int n_fft);
int gl_warps_per_cta(int n_fft);
int gl_ctas_per_sm(int n_fft);
int gl_resident_warps_per_sm(int n_fft);

}  // namespace xdtts
