// Staging of PAGEABLE caller buffers.  The reference hands the vocoder ordinary heap arrays (ndarray / Vec,
// /root/reference src/lib.rs:141-155); the DMA engines need pinned memory, so such buffers go through a pinned
// staging area.  One thread doing that with memcpy costs more than the whole vocode of a 32-utterance batch
// (~15 ms measured in round 1, against 5 ms of kernels); here the copies are cut into chunks, spread over a few
// persistent helper threads and overlapped with the DMA (an utterance is on the bus while the next is being copied).
#pragma once
#include <atomic>
#include <cstddef>

namespace xdtts {

// memcpy(dst, src, bytes) on the helper threads, in chunks; *pending is incremented once per chunk before the call
// returns and decremented as chunks complete (the caller waits for it to reach zero)
// to_staging: dst is a pinned buffer a DMA will read next -> non-temporal stores (see host_copy.cu)
void host_copy_async(void* dst, const void* src, size_t bytes, std::atomic<int>* pending, bool to_staging = false);
// wait until *pending == 0 (the calling thread helps with queued chunks meanwhile)
void host_copy_wait(std::atomic<int>* pending);
int host_copy_threads();

}  // namespace xdtts
