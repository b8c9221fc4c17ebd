// microbench.cu — access-pattern and FMA-throughput probes behind the band-limited kernel design.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static const int H = 240;
static const long long NC = 240LL * 155;

// thread per column, loop over h (pairs h, H-h like k_bl_fwd_h); VEC consecutive columns per thread
template <int VEC, int U>
__global__ void k_colread(const float* __restrict__ x, float* __restrict__ out, long long nc, int n_cblocks) {
    const long long vol = blockIdx.x / n_cblocks;
    const long long c = ((long long)(blockIdx.x - vol * n_cblocks) * blockDim.x + threadIdx.x) * VEC;
    if (c >= nc) return;
    const float* xv = x + vol * H * nc + c;
    float acc[VEC];
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int h = 1; h + U - 1 <= 119; h += U) {
        float a[U][VEC], b[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (VEC == 1) { a[u][0] = __ldcs(xv + (long long)(h + u) * nc); b[u][0] = __ldcs(xv + (long long)(H - h - u) * nc); }
            else if (VEC == 2) { float2 t = __ldcs((const float2*)(xv + (long long)(h + u) * nc)); a[u][0] = t.x; a[u][1] = t.y;
                                 float2 s = __ldcs((const float2*)(xv + (long long)(H - h - u) * nc)); b[u][0] = s.x; b[u][1] = s.y; }
            else { float4 t = __ldcs((const float4*)(xv + (long long)(h + u) * nc)); a[u][0] = t.x; a[u][1] = t.y; a[u][2] = t.z; a[u][3] = t.w;
                   float4 s = __ldcs((const float4*)(xv + (long long)(H - h - u) * nc)); b[u][0] = s.x; b[u][1] = s.y; b[u][2] = s.z; b[u][3] = s.w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            for (int v = 0; v < VEC; ++v) acc[v] += a[u][v] * 1.0001f + b[u][v];
    }
    float s = 0.f;
    for (int v = 0; v < VEC; ++v) s += acc[v];
    out[vol * nc + c] = s;
}

template <int VEC>
__global__ void k_colwrite(float* __restrict__ y, long long nc, int n_cblocks) {
    const long long vol = blockIdx.x / n_cblocks;
    const long long c = ((long long)(blockIdx.x - vol * n_cblocks) * blockDim.x + threadIdx.x) * VEC;
    if (c >= nc) return;
    float* yv = y + vol * H * nc + c;
    for (int h = 1; h <= 119; ++h) {
        float v = (float)h;
        if (VEC == 1) { __stcs(yv + (long long)h * nc, v); __stcs(yv + (long long)(H - h) * nc, v); }
        else if (VEC == 2) { __stcs((float2*)(yv + (long long)h * nc), make_float2(v, v)); __stcs((float2*)(yv + (long long)(H - h) * nc), make_float2(v, v)); }
        else { __stcs((float4*)(yv + (long long)h * nc), make_float4(v, v, v, v)); __stcs((float4*)(yv + (long long)(H - h) * nc), make_float4(v, v, v, v)); }
    }
}

// FMA throughput: NCH independent chains per thread
template <int NCH, bool PACKED>
__global__ void k_fma(float* out, int iters) {
    float2 acc[NCH];
    for (int i = 0; i < NCH; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, 1.f);
    const float2 m = make_float2(1.0001f, 0.9999f), a = make_float2(1e-6f, -1e-6f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (PACKED) acc[i] = __ffma2_rn(acc[i], m, a);
            else { acc[i].x = fmaf(acc[i].x, m.x, a.x); acc[i].y = fmaf(acc[i].y, m.y, a.y); }
        }
    }
    float s = 0.f;
    for (int i = 0; i < NCH; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    const int nvol = 32;
    const size_t n = (size_t)nvol * H * NC;
    float *x, *y, *o;
    CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 4)); CK(cudaMalloc(&o, (size_t)nvol * NC * 4 + (1 << 24)));
    CK(cudaMemset(x, 0, n * 4));
    const double gb = (double)n * 4 / 1e9;
    printf("copy (cudaMemcpy D2D) %.1f GB/s (read+write)\n", 2 * gb / (time_ms([&] { cudaMemcpyAsync(y, x, n * 4, cudaMemcpyDeviceToDevice); }) * 1e-3));
#define RUNREAD(VEC, U, T) { int ncb = (int)((NC / VEC + T - 1) / T); float ms = time_ms([&] { k_colread<VEC, U><<<ncb * nvol, T>>>(x, o, NC, ncb); }); \
        printf("colread  vec=%d U=%d threads=%4d : %.3f ms  %.0f GB/s  (%.2f us/vol)\n", VEC, U, T, ms, gb / (ms * 1e-3), ms * 1e3 / nvol); }
    RUNREAD(1, 4, 128) RUNREAD(1, 4, 256) RUNREAD(1, 4, 512) RUNREAD(1, 4, 1024)
    RUNREAD(1, 8, 256) RUNREAD(1, 16, 256)
    RUNREAD(2, 4, 256) RUNREAD(2, 8, 256) RUNREAD(2, 8, 128)
    RUNREAD(4, 4, 256) RUNREAD(4, 8, 256) RUNREAD(4, 8, 128) RUNREAD(4, 4, 64)
#define RUNWRITE(VEC, T) { int ncb = (int)((NC / VEC + T - 1) / T); float ms = time_ms([&] { k_colwrite<VEC><<<ncb * nvol, T>>>(y, NC, ncb); }); \
        printf("colwrite vec=%d threads=%4d : %.3f ms  %.0f GB/s  (%.2f us/vol)\n", VEC, T, ms, gb / (ms * 1e-3), ms * 1e3 / nvol); }
    RUNWRITE(1, 128) RUNWRITE(1, 256) RUNWRITE(1, 512) RUNWRITE(2, 256) RUNWRITE(4, 256) RUNWRITE(4, 128)
    CK(cudaGetLastError());
    // FMA throughput
    const int iters = 4096;
#define RUNFMA(NCH, PACKED, BLK) { float ms = time_ms([&] { k_fma<NCH, PACKED><<<148 * BLK, 256>>>(o, iters); }); \
        double fma = 148.0 * BLK * 256 * (double)iters * NCH * 2; \
        printf("fma chains=%2d packed=%d blocks/SM=%d : %.3f ms  %.1f TFMA/s (%.1f FMA/clk/SM at 1.9 GHz)\n", NCH, PACKED, BLK, ms, fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / 148 / 1.9e9); }
    RUNFMA(4, false, 4) RUNFMA(8, false, 4) RUNFMA(8, false, 8) RUNFMA(4, true, 4) RUNFMA(8, true, 4) RUNFMA(8, true, 8) RUNFMA(2, true, 8) RUNFMA(2, false, 8)
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
