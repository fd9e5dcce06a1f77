// Forward pass (add-compare-select + decision stores) of the K=7 decoder with one codeword per thread (viterbi_core.cuh, the
// library's kernel) against one codeword per lane PAIR (viterbi_core2.cuh, validated on the CPU by tests/test_host_emul.py), on the
// headline launch size: 32 768 codewords x 3096 steps.  Symbol words come from a per-thread generator (no staging, no
// de-puncturing): the question is only whether twice the warps at ~1.35x the instructions beat 1.73 warps per scheduler.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o vit_threads scripts/ubench/vit_threads.cu && ./vit_threads
#include "../../welle.io_b200/csrc/viterbi_core2.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace dabb;

__device__ __forceinline__ uint32_t nextw(uint32_t& s) { s = s * 1664525u + 1013904223u; return s & 0xFEFEFEFEu; }

__global__ void __launch_bounds__(128, 4) fwd1(uint2* dec, int n_cw, int nsteps, uint32_t one)
{
    const int cw = blockIdx.x * 128 + threadIdx.x;
    if (cw >= n_cw) return;
    uint32_t Q[32]; vit_init(Q);
    uint32_t seed = cw * 2654435761u;
    uint2* d = dec + (int64_t)blockIdx.x * nsteps * 128 + threadIdx.x;
    for (int g = 0; g < nsteps / 6; g++) {
        if ((g & 3) == 0) vit_normalize(Q);
        uint32_t w[6], dd[12];
#pragma unroll
        for (int k = 0; k < 6; k++) w[k] = nextw(seed);
        vit_six_steps(Q, w, dd, one);
#pragma unroll
        for (int k = 0; k < 6; k++) d[(int64_t)(6 * g + k) * 128] = make_uint2(dd[2 * k], dd[2 * k + 1]);
    }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) fwd2(uint32_t* dec, int n_cw, int nsteps)
{
    const int t = threadIdx.x, cw = blockIdx.x * 64 + (t >> 1);
    if (cw >= n_cw) return;
    const uint32_t tau = t & 1u;
    const int lane = t & 31;
    uint32_t Q[16]; vit2_init(Q, tau);
    uint32_t seed = cw * 2654435761u;
    uint32_t* d = dec + (int64_t)blockIdx.x * nsteps * 128 + t;
    for (int g = 0; g < nsteps / 6; g++) {
        if ((g & 3) == 0) {
            uint32_t m = vit2_local_min(Q);
            m = vminu16x2(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
#pragma unroll
            for (int r = 0; r < 16; r++) Q[r] -= m;
        }
        uint32_t w[6], dd[6];
#pragma unroll
        for (int k = 0; k < 6; k++) w[k] = nextw(seed);
        uint32_t X[16], Y[16];
#pragma unroll
        for (int r = 0; r < 16; r++) { X[r] = __shfl_sync(0xFFFFFFFFu, Q[r], lane & ~1); Y[r] = __shfl_sync(0xFFFFFFFFu, Q[r], lane | 1); }
        dd[0] = vit2_acs<0>(Q, X, Y, w[0], tau);
        dd[1] = vit2_acs<1>(Q, X, Y, w[1], tau);
        dd[2] = vit2_acs<2>(Q, X, Y, w[2], tau);
        dd[3] = vit2_acs<3>(Q, X, Y, w[3], tau);
        dd[4] = vit2_acs<4>(Q, X, Y, w[4], tau);
        dd[5] = vit2_acs<5>(Q, X, Y, w[5], tau);
#pragma unroll
        for (int k = 0; k < 6; k++) d[(int64_t)(6 * g + k) * 128] = dd[k];
    }
}

int main()
{
    const int n_cw = 32768, nsteps = 3096;
    uint2* dec; cudaMalloc(&dec, (size_t)n_cw * nsteps * sizeof(uint2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto launch, const char* name) {
        for (int i = 0; i < 3; i++) launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %.4f ms per launch (%s)\n", name, ms / 10, cudaGetErrorString(cudaGetLastError()));
    };
    timeit([&] { fwd1<<<n_cw / 128, 128>>>(dec, n_cw, nsteps, 1u); }, "one thread per codeword");
    timeit([&] { fwd2<4><<<n_cw / 64, 128>>>((uint32_t*)dec, n_cw, nsteps); }, "two threads, >=4 CTAs/SM");
    timeit([&] { fwd2<6><<<n_cw / 64, 128>>>((uint32_t*)dec, n_cw, nsteps); }, "two threads, >=6 CTAs/SM");
    // equal outputs?  both generators give the same symbol words per codeword: compare the survivor decisions through the position maps
    return 0;
}
