// Micro-benchmark: issue throughput of scalar vs packed FP32 instructions on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 16;

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float a, float b) {
    float x[ILP];
    unsigned long long p[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 0.001f + i; p[i] = (unsigned long long)(threadIdx.x + i) * 0x3f8000013f800001ull; }
    const unsigned long long pa = 0x3f7fff003f7fff00ull, pb = 0x3a0000003a000000ull;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                                            // FFMA reg,reg,reg
            if (MODE == 1) x[i] = fmaf(x[i], 0.999f, b);                                       // FFMA imm
            if (MODE == 2) x[i] = x[i] + a;                                                    // FADD
            if (MODE == 3) x[i] = x[i] * a;                                                    // FMUL
            if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));   // FFMA2 regs
            if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));                // FADD2
            if (MODE == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));                // FMUL2
        }
    }
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += x[i] + (float)(p[i] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d, int sms) {
    const int blocks = sms * 8;
    bench<MODE><<<blocks, 256>>>(d, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, 256>>>(d, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 8 * ITERS * ILP;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-18s %8.3f ms  %6.2f warp-instr/clk/SM  (%5.2f per SMSP)\n", name, ms, warp_instr / cycles / sms, warp_instr / cycles / sms / 4);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d;
    cudaMalloc(&d, sizeof(float) * sms * 8 * 256);
    run<0>("FFMA r,r,r", d, sms);
    run<1>("FFMA r,imm,r", d, sms);
    run<2>("FADD", d, sms);
    run<3>("FMUL", d, sms);
    run<4>("FFMA2", d, sms);
    run<5>("FADD2", d, sms);
    run<6>("FMUL2", d, sms);
    return 0;
}
