// tc_rate.cu — how fast does tcgen05.mma kind::tf32 go for the skinny shapes of the pruned DFT?
// `niss` warps of one CTA each issue R MMAs (M, N, K = 8) into their own accumulator (A from shared memory or TMEM);
// the grid is 1 CTA, or enough CTAs for two per SM.  Reports cycles per MMA seen by one issuing warp.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#include "../medical-vision-textural-bias_b200/mvtb/csrc/tc_common.cuh"
using namespace mvtb;

__global__ void __launch_bounds__(192, 2) k_rate(int M, int N, int niss, int ts, int R, uint32_t cols, long long* out, int batch, int nacc, int commit_each, int sync_each) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar[4];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 8 + 256 * 8); i += blockDim.x) ((float*)smem)[i] = 1.0f;
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&s_tmem), cols);
    if (tid == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(tc::smem_u32(&s_bar[i]), i == 3 ? 1000000 : 1); tc::mbar_init_fence(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    if (warp >= 1 && warp <= niss) {
        const int w = warp - 1;
        const uint32_t idesc = tc::idesc_tf32(M, N);
        const uint64_t da = tc::smem_desc(tc::smem_u32(smem), 16 * 128, 128);
        const uint64_t db = tc::smem_desc(tc::smem_u32(smem + 128 * 8 * 4), (uint32_t)(N / 8) * 128, 128);
        const uint32_t a_t = tmem + cols - 16;
        const uint32_t d = tmem + (uint32_t)w * 6u * (uint32_t)N;
        const long long t0 = clock64();
        // `batch` MMAs per elected block (round-robin over `nacc` accumulators inside the block), optional commit per block
        for (int r = 0; r < R; r += batch) {
            if (tc::elect_one()) {
                for (int b = 0; b < batch; ++b) {
                    const uint32_t dd = d + (uint32_t)(b % nacc) * (uint32_t)N;
                    if (ts) tc::mma_ts(dd, a_t, db, idesc, 1);
                    else tc::mma_ss(dd, da, db, idesc, 1);
                }
                if (commit_each) tc::mma_commit(tc::smem_u32(&s_bar[3]));
            }
            if (sync_each) __syncwarp();
        }
        if (tc::elect_one()) tc::mma_commit(tc::smem_u32(&s_bar[w]));
        __syncwarp();
        tc::mbar_wait(tc::smem_u32(&s_bar[w]), 0);
        const long long t2 = clock64();
        if ((tid & 31) == 0 && blockIdx.x == 0 && w == 0) out[0] = t2 - t0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, cols);
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
    const int R = 4092;
    printf("mode  N  issuing_warps  batch nacc commit sync  cyc/mma(per warp)\n");
    for (int ts = 1; ts < 2; ++ts)
        for (int N : {32})
            for (int niss : {1, 2})
                for (int batch : {1, 2, 6})
                    for (int nacc : {1, 6})
                        for (int commit_each : {0, 1})
                            for (int sync_each : {0, 1}) {
                                if (nacc > batch) continue;
                                k_rate<<<1, 192, 32 * 1024>>>(128, N, niss, ts, R - R % batch, 512, d, batch, nacc, commit_each, sync_each);
                                CK(cudaDeviceSynchronize());
                                long long h;
                                CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
                                printf("%s %4d %6d %8d %4d %5d %5d %14.1f\n", ts ? "TS" : "SS", N, niss, batch, nacc, commit_each, sync_each, (double)h / (R - R % batch));
                            }
    return 0;
}
