// k_sync.cu — device-side completion flags for the fused sort-first gather (one process per GPU).
//
// In peer-store mode the ranks' k_tile kernels store their pixels straight into rank 0's framebuffer over
// NVLink (draw_canvas_bind_external on a CUDA IPC pointer).  What is left of the "gather" is knowing when those
// stores have landed — without the host: every rank has a flag word in a small buffer that rank 0 owns and
// exports through CUDA IPC.  After its frame a peer enqueues k_flag_signal (a system-scope release store of the
// frame's sequence number into its word, over NVLink); rank 0 enqueues k_flags_wait on its canvas' stream, which
// polls its own memory until every word has reached the sequence number.  Rank 0 publishes "frame consumed" the
// same way in a word the peers poll before they overwrite the framebuffer with the next frame.  The wait gives up
// after a few seconds (a peer died) and raises an error word instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>

namespace drawb200 {

__global__ void k_flag_signal(uint32_t *flag, uint32_t value) {
    __threadfence_system(); // the frame's stores (previous kernels of this stream) are performed before the flag
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void k_flags_wait(const uint32_t *flags, uint32_t n, uint32_t value, uint32_t *error_word) {
    const uint32_t i = threadIdx.x;
    if (i >= n) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        if ((int32_t)(v - value) >= 0) break; // sequence numbers only grow (wrap-safe compare)
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 5000000000ull) {
            if (error_word) atomicExch(error_word, 1u);
            break;
        }
        __nanosleep(200);
    }
}

cudaError_t launch_flag_signal(uint32_t *flag, uint32_t value, cudaStream_t stream) {
    k_flag_signal<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}
cudaError_t launch_flags_wait(const uint32_t *flags, uint32_t n, uint32_t value, uint32_t *error_word, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_flags_wait<<<1, n <= 32 ? 32 : 64, 0, stream>>>(flags, n, value, error_word);
    return cudaGetLastError();
}

} // namespace drawb200
