// Which pipe does IDP.4A share?  Throughput of PRMT / IDP4A / FFMA streams alone and interleaved on one SM's worth of warps.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(uint32_t *out, uint32_t seed, float fs) {
    uint32_t a[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 8 + i; f[i] = fs + i; }
    const uint32_t bias = seed ^ 0x47000000u, sel = seed | 0xffu;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 3 || MODE == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(a[i]) : "r"(bias));
            if (MODE == 1 || MODE == 3 || MODE == 4) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(sel), "r"(bias));
            if (MODE == 2 || MODE == 4 || MODE == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fs), "f"(fs));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i] + __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, int per_iter) {
    uint32_t *out;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 1024>>>(out, 1, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148, 1024>>>(out, 1, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // warp-instructions per SMSP per ns
    const double winstr = (double)ITERS * 8 * per_iter * (1024 / 32) / 4;  // per SMSP
    printf("%-16s %.3f ms  %.3f warp-instr/ns/SMSP (%.2f per clk at 1.965 GHz)\n", name, ms, winstr / (ms * 1e6), winstr / (ms * 1e6) / 1.965);
    cudaFree(out);
}
int main() {
    run<0>("PRMT", 1); run<1>("IDP4A", 1); run<2>("FFMA", 1); run<3>("PRMT+IDP4A", 2); run<4>("IDP4A+FFMA", 2); run<5>("PRMT+FFMA", 2);
    return 0;
}
