// FP32 FMA peak on B200: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), register-only loops.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu && ./fma_peak
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2 (unsigned long long &acc, unsigned long long a, unsigned long long b)
{ asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
template <int MODE> __global__ void __launch_bounds__ (256) k (float *out, int iters, float a0, float b0)
{
    float acc[32]; unsigned long long acc2[16];
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
    for (int i = 0; i < 16; ++i) acc2[i] = ((unsigned long long) __float_as_uint (acc[2 * i + 1]) << 32) | __float_as_uint (acc[2 * i]);
    float a = a0, b = b0;
    unsigned long long a2 = ((unsigned long long) __float_as_uint (a) << 32) | __float_as_uint (a * 1.0001f);
    unsigned long long b2 = ((unsigned long long) __float_as_uint (b) << 32) | __float_as_uint (b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = fmaf (acc[i], a, b);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) ffma2 (acc2[i], a2, b2);
        }
    }
    float s = 0;
    if (MODE == 0) for (int i = 0; i < 32; ++i) s += acc[i];
    else for (int i = 0; i < 16; ++i) s += __uint_as_float ((unsigned) acc2[i]) + __uint_as_float ((unsigned) (acc2[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main ()
{
    float *out; cudaMalloc (&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
    const int iters = 20000;
    for (int warpsPerSm : {8, 16, 32, 64}) {
        const int blocks = 148 * warpsPerSm / 8;
        for (int mode = 0; mode < 2; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord (e0);
                if (mode == 0) k<0><<<blocks, 256>>> (out, iters, 0.999f, 0.5f); else k<1><<<blocks, 256>>> (out, iters, 0.999f, 0.5f);
                cudaEventRecord (e1); cudaEventSynchronize (e1);
            }
            float ms; cudaEventElapsedTime (&ms, e0, e1);
            const double fmas = (double) blocks * 256 * iters * 128.0;
            printf ("%2d warps/SM %-6s: %7.2f TFMA/s  (%.2f FMA/clk/SM at 1.965 GHz)\n", warpsPerSm, mode ? "FFMA2" : "FFMA", fmas / ms / 1e9,
                    fmas / ms / 1e3 / 148 / 1.965e6);
        }
    }
    return 0;
}
