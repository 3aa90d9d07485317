// fusg_common.h -- shared host-side helpers of the C-ABI library (error capture, launch counter).
#pragma once
#include <cuda_runtime.h>

// records cudaGetLastError() text; returns FUSG_OK or FUSG_ERR_CUDA
int fusg_check_launch();
int fusg_record_cuda(cudaError_t e);
void fusg_count_launch(int n);

// Per-device lazily initialised state.  cudaFuncSetAttribute, SM counts and helper streams belong to ONE device; a
// process may drive several (and from several host threads), so everything of that kind is keyed by the current
// device and guarded by a mutex instead of living in a function-local static.
constexpr int FUSG_MAX_DEVICES = 64;
int fusg_current_device();                     // cudaGetDevice(), clamped into [0, FUSG_MAX_DEVICES)
int fusg_num_sms();                            // multiprocessor count of the current device (cached per device)
// Runs `fn()` once per (device, slot); returns its cudaError_t (cudaSuccess on later calls).  slot < 16.
// A slot that carries a size (dynamic shared memory limit) is re-run when `size` exceeds what was set before.
#include <functional>
cudaError_t fusg_once_per_device(int slot, size_t size, const std::function<cudaError_t()> &fn);

#ifdef __CUDACC__
// Bounded spin for mbarrier waits: a wait that has not completed after ~4 s of wall time (a descriptor / pipeline bug --
// healthy waits take microseconds) aborts the kernel with a device-side assertion instead of hanging the GPU; the
// next library call then returns FUSG_ERR_CUDA and fusg_last_error() names cudaErrorAssert.
extern "C" __device__ void __assertfail(const char *message, const char *file, unsigned line, const char *function, size_t charSize);
__device__ __forceinline__ void fusg_spin_guard(unsigned long long &t0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) __assertfail("fusg: mbarrier wait timed out (TMA descriptor / pipeline bug)", __FILE__, __LINE__, "mbar_wait", 1);
}
#endif
