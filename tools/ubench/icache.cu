// Instruction-cache capacity probe for sm_100a: every warp loops over a straight-line body of K FFMAs (16 B each).
// Reports SM cycles per warp-instruction for body sizes 4 KB .. 512 KB with 8 one-warp CTAs per SM (the solver's shape),
// (a) all warps running the same body, (b) the 8 warps of an SM running 8 different bodies (disjoint code).
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__device__ __forceinline__ void body(float& a0, float& a1, float& a2, float& a3, float& a4, float& a5, float& a6, float& a7, float m, float c)
{
#pragma unroll
    for (int i = 0; i < K / 8; i++) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a4) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a5) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a6) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a7) : "f"(m), "f"(c));
    }
}

// NV variants of the body are separate code regions (the switch keeps them all alive); variant = blockIdx.x % nv
template <int K, int NV>
__global__ void __launch_bounds__(32) probe(float* out, int iters, int nv, long long* cyc)
{
    float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    const float m = 0.999f, c = 0.001f;
    const int v = (blockIdx.x / 148) % nv;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        switch (v) {
        case 0: body<K>(a0, a1, a2, a3, a4, a5, a6, a7, m, c); break;
        case 1: if (NV > 1) body<K>(a1, a0, a2, a3, a4, a5, a6, a7, m, c); break;
        case 2: if (NV > 1) body<K>(a2, a1, a0, a3, a4, a5, a6, a7, m, c); break;
        case 3: if (NV > 1) body<K>(a3, a1, a2, a0, a4, a5, a6, a7, m, c); break;
        case 4: if (NV > 1) body<K>(a4, a1, a2, a3, a0, a5, a6, a7, m, c); break;
        case 5: if (NV > 1) body<K>(a5, a1, a2, a3, a4, a0, a6, a7, m, c); break;
        case 6: if (NV > 1) body<K>(a6, a1, a2, a3, a4, a5, a0, a7, m, c); break;
        default: if (NV > 1) body<K>(a7, a1, a2, a3, a4, a5, a6, a0, m, c); break;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, int NV>
void run(const char* tag, float* out, long long* cyc, int nv)
{
    const int blocks = 148 * 8;
    const int iters = (1 << 22) / K;             // ~4M instructions per warp
    // pad shared memory so that exactly 8 CTAs fit per SM, as in the solver
    cudaFuncSetAttribute(probe<K, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 27000);
    probe<K, NV><<<blocks, 32, 27000>>>(out, 4, nv, cyc);
    cudaDeviceSynchronize();
    probe<K, NV><<<blocks, 32, 27000>>>(out, iters, nv, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < blocks; i++) s += (double)h[i];
    printf("%s body %4d KB x %d variant(s): %.3f cycles per warp-instruction (%s)\n", tag, K * 16 / 1024, nv, s / blocks / ((double)iters * K),
           cudaGetErrorString(e));
}

int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 8 * 32 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
    run<256, 1>("same", out, cyc, 1);
    run<512, 1>("same", out, cyc, 1);
    run<1024, 1>("same", out, cyc, 1);
    run<2048, 1>("same", out, cyc, 1);
    run<2560, 1>("same", out, cyc, 1);
    run<3072, 1>("same", out, cyc, 1);
    run<3584, 1>("same", out, cyc, 1);
    run<4096, 1>("same", out, cyc, 1);
    run<6144, 1>("same", out, cyc, 1);
    run<8192, 1>("same", out, cyc, 1);
    run<12288, 1>("same", out, cyc, 1);
    run<16384, 1>("same", out, cyc, 1);
    run<256, 8>("diff", out, cyc, 8);
    run<320, 8>("diff", out, cyc, 8);
    run<384, 8>("diff", out, cyc, 8);
    run<448, 8>("diff", out, cyc, 8);
    run<512, 8>("diff", out, cyc, 8);
    run<768, 8>("diff", out, cyc, 8);
    run<1024, 8>("diff", out, cyc, 8);
    run<2048, 8>("diff", out, cyc, 8);
    return 0;
}
