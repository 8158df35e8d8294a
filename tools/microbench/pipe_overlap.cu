// Micro-benchmark: can an SM overlap packed-FP32 arithmetic (FFMA2, the FMA pipe) of some warps with 128-bit
// shared-memory loads (LDS.128, the LSU data pipe) of other warps?  K1 keeps the FMA pipe busy 49 % and the LSU pipe
// 59 % of the time and takes about the SUM of the two; this separates "the hardware cannot overlap them" from "K1's
// instruction streams do not".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_overlap pipe_overlap.cu && ./pipe_overlap
// 16 warps per SM (one 512-thread CTA per SM).  role 0: even warps FFMA2, odd warps idle; role 1: odd warps LDS.128,
// even warps idle; role 2: both at once; role 3: EVERY warp alternates bursts of the two (64 FFMA2, then 8 LDS.128),
// the way a K1 warp moves through its stages.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS_F = 4096, ITERS_L = 512, ILP = 16;

__global__ void __launch_bounds__(512, 1) bench(int role, unsigned* out) {
    __shared__ uint4 s[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long p[ILP];
    for (int i = 0; i < ILP; ++i) p[i] = (unsigned long long)(threadIdx.x + i) * 0x3f8000013f800001ull;
    const unsigned long long pa = 0x3f7fff003f7fff00ull, pb = 0x3a0000003a000000ull;
    unsigned acc = 0;
    const bool fma_warp = role == 3 || (role != 1 && (warp & 1) == 0);
    const bool lds_warp = role == 3 || (role != 0 && (warp & 1) == 1);
    if (role == 3) {
        // same total work per SM as role 2: every warp does half of each kind, in alternating bursts
        for (int it = 0; it < ITERS_L / 2; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                const uint4 v = s[(lane + (it * ILP + i) * 32) & 2047];
                acc ^= v.x ^ v.y;
                acc ^= v.z ^ v.w;
            }
        }
    } else {
        if (fma_warp)
            for (int it = 0; it < ITERS_F; ++it)
#pragma unroll
                for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        if (lds_warp)
            for (int it = 0; it < ITERS_L; ++it)
#pragma unroll
                for (int i = 0; i < ILP; ++i) {
                    const uint4 v = s[(lane + (it * ILP + i) * 32) & 2047];
                    acc ^= v.x ^ v.y;
                    acc ^= v.z ^ v.w;
                }
    }
    unsigned t = acc;
    for (int i = 0; i < ILP; ++i) t ^= (unsigned)p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    unsigned* d;
    cudaMalloc(&d, sizeof(unsigned) * sms * 512);
    const char* names[4] = {"FFMA2 on 8 warps", "LDS.128 on 8 warps", "both at once (8 + 8 warps)", "16 warps alternating bursts"};
    for (int role = 0; role < 4; ++role) {
        bench<<<sms, 512>>>(role, d);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        bench<<<sms, 512>>>(role, d);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double cycles = ms * 1e-3 * clk * 1e3;
        const double ffma2 = 8.0 * ITERS_F * ILP, lds = 8.0 * ITERS_L * ILP;       // warp-instructions per SM (roles 2, 3)
        printf("%-30s %8.3f ms  %9.0f clk/SM", names[role], ms, cycles);
        if (role == 0) printf("   FFMA2 %.2f warp-instr/clk/SM", ffma2 / cycles);
        if (role == 1) printf("   LDS.128 %.3f warp-instr/clk/SM (%.2f wavefronts/clk)", lds / cycles, 4 * lds / cycles);
        printf("\n");
    }
    return 0;
}
