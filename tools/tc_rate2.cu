// tc_rate2.cu — cost of the MMA issuer's loop body: mma only, mma + commit, mma + commit every 2nd/4th iteration,
// with 1..6 issuing warps (each its own accumulator and barrier set).  Cycles per iteration seen by warp 0.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#include "../medical-vision-textural-bias_b200/mvtb/csrc/tc_common.cuh"
using namespace mvtb;

__global__ void __launch_bounds__(256, 1) k_rate(int N, int niss, int R, int commit_every, int fence, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar[8][8];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 8 + 256 * 8); i += blockDim.x) ((float*)smem)[i] = 1.0f;
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&s_tmem), 512);
    if (tid == 0) { for (int i = 0; i < 64; ++i) tc::mbar_init(tc::smem_u32(&s_bar[i / 8][i % 8]), 1); tc::mbar_init_fence(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (warp >= 1 && warp <= niss) {
        const int w = warp - 1;
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t b0 = tc::smem_u32(smem + 128 * 8 * 4);
        const uint32_t d = tmem + 256 + (uint32_t)w * (uint32_t)N;
        const long long t0 = clock64();
        uint32_t acc = 0;
        for (int r = 0; r < R; ++r) {
            const int sl = r & 7;
            if (fence) tc::fence_after_sync();
            if (tc::elect_one()) {
                tc::mma_ts(d, tmem + (uint32_t)sl * 32u + (uint32_t)(w & 1) * 8u, tc::smem_desc(b0, (uint32_t)(N / 8) * 128u, 128u), idesc, acc);
                if (commit_every && (r % commit_every) == commit_every - 1) tc::mma_commit(tc::smem_u32(&s_bar[w][sl]));
            }
            acc = 1;
            __syncwarp();
        }
        const long long t1 = clock64();
        if (tc::elect_one()) tc::mma_commit(tc::smem_u32(&s_bar[w][7]));
        __syncwarp();
        // the last barrier may have completed several phases; just wait until all MMAs are done via a fresh commit on a spare barrier
        if ((tid & 31) == 0 && w == 0) { out[0] = t1 - t0; }
    }
    __syncthreads();
    // drain: wait a while for the tensor pipe before freeing TMEM
    if (warp == 1) { long long t = clock64(); while (clock64() - t < 2000000) {} }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
    const int R = 4096;
    printf("N  issuing_warps  commit_every  fence  cyc/iteration(warp 0)\n");
    for (int N : {32})
        for (int niss : {1, 3, 6})
            for (int ce : {0, 1, 2, 4})
                for (int fence : {0, 1}) {
                    k_rate<<<1, 256, 32 * 1024>>>(N, niss, R, ce, fence, d);
                    CK(cudaDeviceSynchronize());
                    long long h;
                    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
                    printf("%d %6d %10d %8d %14.1f\n", N, niss, ce, fence, (double)h / R);
                }
    return 0;
}
