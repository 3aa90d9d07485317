// common.cu -- library-wide state of libfusg.so: last CUDA error text and the launch counter
// that bench.py reports as "gpu_launches".
#include <atomic>
#include <cstdio>
#include <cstring>
#include "../../include/fusg.h"
#include "fusg_common.h"

static char g_last_error[256] = "";
static std::atomic<int> g_launches{0};

int fusg_record_cuda(cudaError_t e) {
    if (e == cudaSuccess) return FUSG_OK;
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return FUSG_ERR_CUDA;
}

int fusg_check_launch() { return fusg_record_cuda(cudaGetLastError()); }

void fusg_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int fusg_version(void) { return 100; }
extern "C" const char *fusg_last_error(void) { return g_last_error; }
extern "C" int fusg_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
