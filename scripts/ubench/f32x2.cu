// micro-benchmark: issue/pipe throughput of packed FP32 (FADD2/FMUL2/FFMA2) against scalar FADD/FMUL/FFMA on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f32x2 f32x2.cu ; run: ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mul1(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE> __global__ void k(float* out, int iters, float seed)
{
    constexpr int N = 8;   // independent chains per thread
    if (MODE == 0 || MODE == 1) {           // scalar: 16 floats
        float v[2 * N];
        for (int i = 0; i < 2 * N; i++) v[i] = seed + i + threadIdx.x;
        const float c = seed * 0.5f + 1.0f;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 2 * N; i++) v[i] = MODE == 0 ? add1(v[i], c) : mul1(v[i], c);
        float s = 0; for (int i = 0; i < 2 * N; i++) s += v[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        u64 v[N];
        for (int i = 0; i < N; i++) { float2 f = make_float2(seed + i + threadIdx.x, seed - i); v[i] = *reinterpret_cast<u64*>(&f); }
        float2 cf = make_float2(seed * 0.5f + 1.0f, seed * 0.25f + 1.0f); const u64 c = *reinterpret_cast<u64*>(&cf);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = MODE == 2 ? add2(v[i], c) : MODE == 3 ? mul2(v[i], c) : fma2(v[i], c, c);
        float s = 0; for (int i = 0; i < N; i++) { float2 f = *reinterpret_cast<float2*>(&v[i]); s += f.x + f.y; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}
template <int MODE> void run(const char* name, int flops_per_instr, int instr_per_iter)
{
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 20000;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    k<MODE><<<blocks, threads>>>(out, 100, 1.0f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, threads>>>(out, iters, 1.0f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double warp_instr = (double)blocks * threads / 32 * iters * instr_per_iter;
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-8s %8.3f ms  %.2f warp-instr/clk/SM  %.1f lane-flop/clk/SM (at %d MHz nominal)\n", name, ms, warp_instr / cyc / p.multiProcessorCount,
           warp_instr * 32 * flops_per_instr / cyc / p.multiProcessorCount, clk / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("FADD", 1, 16); run<1>("FMUL", 1, 16); run<2>("FADD2", 2, 8); run<3>("FMUL2", 2, 8); run<4>("FFMA2", 4, 8);
    return 0;
}
