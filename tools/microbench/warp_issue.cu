// Micro-benchmark: FP32 arithmetic throughput of an SM as a function of the number of resident warps, scalar
// (FFMA) against packed (FFMA2, two FMAs per lane), 16 independent accumulators per thread.  K1 runs 16 warps per SM
// (4 per scheduler), half of them in a transform phase at any time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_issue warp_issue.cu && ./warp_issue
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 8192, ILP = 16;

template <int MODE>
__global__ void bench(float* out, float a, float b) {
    float x[ILP];
    unsigned long long p[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 0.001f + i; p[i] = (unsigned long long)(threadIdx.x + i) * 0x3f8000013f800001ull; }
    const unsigned long long pa = 0x3f7fff003f7fff00ull, pb = 0x3a0000003a000000ull;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        }
    }
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += x[i] + (float)(p[i] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d, int sms, int warps, int clk) {
    bench<MODE><<<sms, warps * 32>>>(d, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<sms, warps * 32>>>(d, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double wi = (double)warps * ITERS * ILP;
    printf("%-6s %2d warps/SM: %5.2f warp-instr/clk/SM = %5.1f FMA/clk/SM, %5.3f instr/clk per warp\n", name, warps, wi / cycles,
           wi / cycles * 32 * (MODE == 1 ? 2 : 1), wi / cycles / warps);
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* d;
    cudaMalloc(&d, sizeof(float) * sms * 1024);
    for (int w : {4, 8, 16, 32}) { run<0>("FFMA", d, sms, w, clk); run<1>("FFMA2", d, sms, w, clk); }
    return 0;
}
