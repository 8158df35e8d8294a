// Library-level entry points: version, thread-local error text, device info cache.
#include "dam_common.cuh"

#include <atomic>
#include <mutex>
#include <string>

static thread_local std::string g_last_error;

void dam_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

void dam_clear_error() { g_last_error.clear(); }

static std::atomic<long long> g_launches{0};
void dam_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t dam_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern const char* dam_k1_last_kernel_name;          // stft_db.cu / stft_tma.cu
extern "C" const char* dam_last_stft_kernel(void) { return dam_k1_last_kernel_name; }

int dam_sm_count() {
    static std::mutex mu;
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(mu);
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

extern "C" const char* dam_version(void) { return "0.1.0"; }

extern "C" const char* dam_last_error(void) { return g_last_error.c_str(); }

extern "C" int64_t dam_stft_num_frames(int64_t n_samples, int hop) {
    if (hop <= 0 || n_samples < 0) return -1;
    return 1 + n_samples / hop;
}
