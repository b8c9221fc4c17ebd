// l2probe.cu — how much freshly WRITTEN data does the B200 L2 keep for a later sparse read-modify-write?
//
// The fused inverse + select kernel (bandlimited_sp.cuh) bets on the select pass finding the inverse pass's output
// lines still in L2.  This probe measures that directly: kernel `wr` writes a buffer of S bytes with coalesced 8-byte
// stores (plain / streaming / L2 evict_last), kernel `touch` then stores 4 bytes into about a third of its 32-byte
// sectors (the select pass's pattern at p = 0.05).  Timed with CUDA events here; run it under
//   ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
// to see the DRAM bytes of every launch.   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2probe l2probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int POL>
__global__ void wr(float2* __restrict__ p, size_t n2, float v) {
    unsigned long long pol = 0;
    if (POL == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const float2 val = make_float2(v + (float)i, v);
        if (POL == 0) p[i] = val;
        else if (POL == 1) __stcs(p + i, val);
        else asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p + i), "f"(val.x), "f"(val.y), "l"(pol) : "memory");
    }
}

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// one thread per 32-byte sector; a third of the sectors get one 4-byte store
__global__ void touch(float* __restrict__ p, size_t nsect, unsigned salt) {
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < nsect; s += (size_t)gridDim.x * blockDim.x) {
        const unsigned h = hash32((unsigned)s ^ salt);
        if (h % 3u == 0u) p[s * 8 + (h >> 29)] = -1.f;
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char** argv) {
    const size_t big = (size_t)1 << 30;
    float* buf;
    CK(cudaMalloc(&buf, big));
    float* flush;
    CK(cudaMalloc(&flush, big));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const int grid = 148 * 8;
    if (argc > 1) {   // short sequence for ncu: flush, wr, touch per (policy, size); read the DRAM bytes per launch
        for (int pol : {0, 2})
            for (int mb : {16, 32, 64, 96}) {
                const size_t bytes = (size_t)mb << 20;
                wr<0><<<grid, 256>>>((float2*)flush, big / 8, 1.f);
                if (pol == 0) wr<0><<<grid, 256>>>((float2*)buf, bytes / 8, 2.f);
                else wr<2><<<grid, 256>>>((float2*)buf, bytes / 8, 2.f);
                touch<<<grid, 256>>>(buf, bytes / 32, mb);
                CK(cudaDeviceSynchronize());
            }
        return 0;
    }
    const int sizes_mb[] = {8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 192, 256};
    printf("policy size_MB  wr_GBps  touch_us  touch_us_cold  (touch after a flush of L2 = cold)\n");
    for (int pol = 0; pol < 3; ++pol) {
        for (int si = 0; si < (int)(sizeof(sizes_mb) / sizeof(int)); ++si) {
            const size_t bytes = (size_t)sizes_mb[si] << 20;
            const size_t n2 = bytes / 8, nsect = bytes / 32;
            float best_wr = 1e9f, best_touch = 1e9f, best_cold = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                wr<0><<<grid, 256>>>((float2*)flush, big / 8, 1.f);             // flush L2 with other data
                CK(cudaEventRecord(e0));
                if (pol == 0) wr<0><<<grid, 256>>>((float2*)buf, n2, 2.f);
                else if (pol == 1) wr<1><<<grid, 256>>>((float2*)buf, n2, 2.f);
                else wr<2><<<grid, 256>>>((float2*)buf, n2, 2.f);
                CK(cudaEventRecord(e1));
                touch<<<grid, 256>>>(buf, nsect, rep);
                CK(cudaEventRecord(e2));
                CK(cudaDeviceSynchronize());
                const float twr = time_ms(e0, e1), tt = time_ms(e1, e2);
                if (twr < best_wr) best_wr = twr;
                if (tt < best_touch) best_touch = tt;
                // cold: flush, then touch
                wr<0><<<grid, 256>>>((float2*)flush, big / 8, 1.f);
                CK(cudaEventRecord(e1));
                touch<<<grid, 256>>>(buf, nsect, rep + 100);
                CK(cudaEventRecord(e2));
                CK(cudaDeviceSynchronize());
                const float tc = time_ms(e1, e2);
                if (tc < best_cold) best_cold = tc;
            }
            printf("%d %6d %8.0f %9.1f %9.1f\n", pol, sizes_mb[si], bytes / best_wr / 1e6, best_touch * 1e3, best_cold * 1e3);
        }
    }
    // pipelined like the fused kernel: write chunk k, touch chunk k-1 (reuse distance = one to two chunks)
    printf("pipelined: chunk_MB  us_per_chunk(wr+touch)  GBps_written\n");
    for (int mb : {16, 32, 36, 48, 64}) {
        const size_t bytes = (size_t)mb << 20;
        const int nch = (int)(big / bytes) < 24 ? (int)(big / bytes) : 24;
        CK(cudaEventRecord(e0));
        for (int k = 0; k <= nch; ++k) {
            if (k < nch) wr<0><<<grid, 256>>>((float2*)((char*)buf + (size_t)k * bytes), bytes / 8, 3.f);
            if (k > 0) touch<<<grid, 256>>>((float*)((char*)buf + (size_t)(k - 1) * bytes), bytes / 32, k);
        }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        const float ms = time_ms(e0, e1);
        printf("%d %9.1f %9.0f\n", mb, ms * 1e3 / nch, (double)bytes * nch / ms / 1e6);
    }
    return 0;
}
