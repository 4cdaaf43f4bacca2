// MUFU.EX2 issue rate per SM sub-partition as a function of resident warps and of interleaved FMA-pipe work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mufu_rate tools/microbench/mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int kFma>
__global__ void k(float* out, int iters, float seed) {
    float x[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; acc[i] = 0.f; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
#pragma unroll
            for (int f = 0; f < kFma; ++f) acc[i] = fmaf(acc[i], 0.999f, y);
            if (kFma == 0) acc[i] += y;
            x[i] = y * 0.5f - 1.0f;
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 2)[0] = t1 - t0;
}

template <int kFma>
void run(float* d, int warps_per_smsp) {
    const int iters = 2000;
    k<kFma><<<148, warps_per_smsp * 128>>>(d, iters, -0.3f);
    cudaDeviceSynchronize();
    k<kFma><<<148, warps_per_smsp * 128>>>(d, iters, -0.3f);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, d + 2, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = double(iters) * 8 * warps_per_smsp;  // MUFU warp-instructions per sub-partition
    printf("fma_per_mufu=%d warps/SMSP=%d : %.2f clk per MUFU warp-instr per SMSP (%.2f clk per warp)\n", kFma + (kFma == 0),
           warps_per_smsp, cyc / per_smsp, cyc / (double(iters) * 8));
}

int main() {
    float* d;
    cudaMalloc(&d, 64);
    for (int w : {1, 2, 4, 8}) run<0>(d, w);
    for (int w : {1, 2, 4}) run<2>(d, w);
    for (int w : {1, 2, 4}) run<4>(d, w);
    return 0;
}
