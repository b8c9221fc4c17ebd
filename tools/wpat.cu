// wpat.cu — how fast can a [vol][H][NC] float batch be WRITTEN when each CTA owns a tile of `seg` columns x all H rows
// (the inverse H pass's store pattern), against a linear fill?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// tile = seg columns; thread t of the CTA writes `vec` floats at column c0 + vec * (t % (seg / vec)), rows r0 + t / (seg / vec) step rows_par
template <int VEC, int CS>
__global__ void __launch_bounds__(512) k_tiles(float* out, int H, int NC, int nvol, int seg, int rows_per_pass) {
    const int tiles_per_vol = (NC + seg - 1) / seg;
    const long long n_tiles = (long long)tiles_per_vol * nvol;
    long long t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x, t_step = 1;
    if (rows_per_pass == 1) { t_lo = blockIdx.x; t_hi = n_tiles; t_step = gridDim.x; }          // strided: the CTAs write adjacent tiles at the same time
    if (rows_per_pass >= 2) {                                                                     // P volumes at a time, strided within each
        const int P = rows_per_pass, C = gridDim.x / P, c = blockIdx.x % C, p = blockIdx.x / C;
        const int tpr = seg / VEC, rpp = blockDim.x / tpr, tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
        if (p >= P) return;
        for (int vol = p; vol < nvol; vol += P)
            for (int tv = c; tv < tiles_per_vol; tv += C) {
                const int col = tv * seg + tc * VEC;
                if (col >= NC || tr >= rpp) continue;
                float* q = out + ((size_t)vol * H + tr) * NC + col;
                for (int r = tr; r < H; r += rpp, q += (size_t)rpp * NC) {
                    if (VEC == 1) { if (CS) __stcs(q, 1.f); else *q = 1.f; }
                    else { float4 v = make_float4(1.f, 2.f, 3.f, 4.f); if (CS) __stcs((float4*)q, v); else *(float4*)q = v; }
                }
            }
        return;
    }
    const int tpr = seg / VEC;                              // threads per row segment
    const int rpp = blockDim.x / tpr;                       // rows written per pass by the CTA
    const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
    (void)rows_per_pass;
    for (long long t = t_lo; t < t_hi; t += t_step) {
        const int vol = (int)(t / tiles_per_vol), c0 = (int)(t - (long long)vol * tiles_per_vol) * seg;
        const int col = c0 + tc * VEC;
        if (col >= NC || tr >= rpp) continue;
        float* p = out + ((size_t)vol * H + tr) * NC + col;
        for (int r = tr; r < H; r += rpp, p += (size_t)rpp * NC) {
            if (VEC == 1) { if (CS) __stcs(p, 1.f); else *p = 1.f; }
            else { float4 v = make_float4(1.f, 2.f, 3.f, 4.f); if (CS) __stcs((float4*)p, v); else *(float4*)p = v; }
        }
    }
}

__global__ void k_linear(float4* out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(out + i, make_float4(1.f, 2.f, 3.f, 4.f));
}

int main() {
    const int H = 240, NC = 37200, nvol = 64;
    const size_t n = (size_t)nvol * H * NC;
    float* d;
    CK(cudaMalloc(&d, n * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto launch, const char* name) {
        for (int i = 0; i < 2; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-44s %8.3f ms  %7.1f GB/s  (%5.2f us per volume)\n", name, ms, n * 4 / ms * 1e-6, ms * 1000 / nvol);
        return 0;
    };
    time([&] { k_linear<<<148 * 8, 512>>>((float4*)d, n / 4); }, "linear float4 .cs");
    CK(cudaGetLastError());
    char name[128];
    for (int grid_mul = 1; grid_mul <= 2; ++grid_mul)
        for (int seg : {128, 512, 2048}) {
            snprintf(name, sizeof name, "seg %4d cols, 4 B/thread, .cs, %d CTA/SM", seg, grid_mul);
            if (seg <= 512) time([&] { k_tiles<1, 1><<<148 * grid_mul, 512>>>(d, H, NC, nvol, seg, 0); }, name);
            snprintf(name, sizeof name, "seg %4d cols, 16 B/thread, .cs, %d CTA/SM", seg, grid_mul);
            time([&] { k_tiles<4, 1><<<148 * grid_mul, 512>>>(d, H, NC, nvol, seg, 0); }, name);
            snprintf(name, sizeof name, "seg %4d cols, 16 B/thread, plain, %d CTA/SM", seg, grid_mul);
            time([&] { k_tiles<4, 0><<<148 * grid_mul, 512>>>(d, H, NC, nvol, seg, 0); }, name);
        }
    for (int mode : {1, 2, 4}) {
        snprintf(name, sizeof name, "seg 128, 4 B/thread, .cs, order mode %d", mode);
        time([&] { k_tiles<1, 1><<<148, 512>>>(d, H, NC, nvol, 128, mode); }, name);
        snprintf(name, sizeof name, "seg 128, 16 B/thread, .cs, order mode %d", mode);
        time([&] { k_tiles<4, 1><<<148, 512>>>(d, H, NC, nvol, 128, mode); }, name);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}
