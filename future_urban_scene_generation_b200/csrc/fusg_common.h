// fusg_common.h -- shared host-side helpers of the C-ABI library (error capture, launch counter).
#pragma once
#include <cuda_runtime.h>

// records cudaGetLastError() text; returns FUSG_OK or FUSG_ERR_CUDA
int fusg_check_launch();
int fusg_record_cuda(cudaError_t e);
void fusg_count_launch(int n);
