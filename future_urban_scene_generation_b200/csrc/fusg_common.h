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
// static shared memory (it also sits at the shared-memory cap): every try_wait that comes back empty-handed -- with the
// 10 ms suspend hint the tensor-core kernels use, or after ~1 us of polling in the gather kernel -- bumps one CTA-wide
// 32-bit counter whose shared-memory address the caller derives from the barrier's own address; a healthy pipeline never
// comes near the limit, a TMA-descriptor / pipeline bug reaches it within seconds and makes the kernel trap instead of
// hanging the GPU (the next library call returns FUSG_ERR_CUDA and fusg_last_error() names the launch failure).
// Kernels zero the counter before their first wait.
template <unsigned LIMIT>
__device__ __forceinline__ void fusg_wait_failed(unsigned counter_smem_addr) {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(counter_smem_addr) : "memory");
    if (old > LIMIT) __trap();      // (a trap, not an assert: no call, hence no ABI stack frame in the hot kernels)
}
#endif
