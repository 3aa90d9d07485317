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
// Bounded mbarrier waits without a live register in the hot kernels (k_conv_tc sits exactly at its register cap) and without
// static shared memory (it also sits at the shared-memory cap): the kernel stores its start time (%globaltimer_hi) in a
// 32-bit shared-memory slot before its first wait; every try_wait that comes back empty-handed re-reads the timer and traps
// once the KERNEL has been running for more than ~10 s -- healthy kernels of this library run for milliseconds, a
// TMA-descriptor / pipeline bug would otherwise spin forever and take the GPU with it.  After the trap the next library call
// returns FUSG_ERR_CUDA and fusg_last_error() names the launch failure.  The slot's address is derived by the caller from the
// barrier's own address, so nothing has to stay live across the wait.
// (the upper half of the nanosecond timer ticks every 4.29 s: 32-bit arithmetic, trap after 3 ticks = 8.6 .. 12.9 s)
__device__ __forceinline__ void fusg_wait_guard_start(unsigned *slot) {
    unsigned now;
    asm volatile("mov.u32 %0, %%globaltimer_hi;" : "=r"(now));
    *slot = now;
}
__device__ __forceinline__ void fusg_wait_failed(unsigned slot_smem_addr) {
    unsigned t0, now;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t0) : "r"(slot_smem_addr) : "memory");
    asm volatile("mov.u32 %0, %%globaltimer_hi;" : "=r"(now));
    if (now - t0 >= 3u) __trap();      // (a trap, not an assert: no call, hence no ABI stack frame in the hot kernels)
}
#endif
