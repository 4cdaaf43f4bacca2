// Does cvt.rn.f16x2.f32 (F2FP.PACK_AB) share an issue port with MUFU.EX2?  Per iteration: 8 ex2 (+ kCvt packs).
#include <cstdio>
#include <cuda_runtime.h>

template <int kCvt, int kMufu>
__global__ void k(float* out, int iters, float seed) {
    float x[8];
    unsigned accu = 0;
    float accf = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kMufu) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[i]) : "f"(x[i]));
            else y[i] = fmaf(x[i], 0.5f, 0.25f);
            accf += y[i];
        }
#pragma unroll
        for (int i = 0; i < kCvt; ++i) {
            unsigned h;
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[(2 * i) & 7]), "f"(y[(2 * i + 1) & 7]));
            accu ^= h;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = y[i] * 0.5f - 1.0f;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = accf + accu;
        reinterpret_cast<long long*>(out + 2)[0] = t1 - t0;
    }
}

template <int kCvt, int kMufu>
void run(float* d, int warps_per_smsp) {
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        k<kCvt, kMufu><<<148, warps_per_smsp * 128>>>(d, iters, -0.3f);
        cudaDeviceSynchronize();
    }
    long long cyc;
    cudaMemcpy(&cyc, d + 2, 8, cudaMemcpyDeviceToHost);
    printf("mufu=%d cvt_per_8=%d warps/SMSP=%d : %.2f clk per iteration per warp (8 ex2 alone = 64)\n", kMufu * 8, kCvt,
           warps_per_smsp, cyc / double(iters));
}

int main() {
    float* d;
    cudaMalloc(&d, 64);
    run<0, 1>(d, 1);
    run<4, 1>(d, 1);
    run<8, 1>(d, 1);
    run<4, 1>(d, 2);
    run<8, 0>(d, 1);
    run<8, 0>(d, 2);
    run<8, 0>(d, 4);
    return 0;
}
