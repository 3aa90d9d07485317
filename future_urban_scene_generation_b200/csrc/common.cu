// common.cu -- library-wide state of libfusg.so: last CUDA error text and the launch counter
// that bench.py reports as "gpu_launches".
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
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

static std::mutex g_dev_mutex;
static int g_num_sms[FUSG_MAX_DEVICES];
static size_t g_once[FUSG_MAX_DEVICES][16];      // 0 = not yet run; otherwise 1 + the size it was run with

int fusg_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FUSG_MAX_DEVICES) dev = 0;
    return dev;
}

int fusg_num_sms() {
    const int dev = fusg_current_device();
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    if (!g_num_sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        g_num_sms[dev] = n > 0 ? n : 148;
    }
    return g_num_sms[dev];
}

cudaError_t fusg_once_per_device(int slot, size_t size, const std::function<cudaError_t()> &fn) {
    const int dev = fusg_current_device();
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    if (g_once[dev][slot] >= 1 + size) return cudaSuccess;
    const cudaError_t e = fn();
    if (e == cudaSuccess) g_once[dev][slot] = 1 + size;
    return e;
}

extern "C" int fusg_version(void) { return 200; }
extern "C" const char *fusg_last_error(void) { return g_last_error; }
extern "C" int fusg_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
