// intensity.cu — the intensity prologue that precedes the k-space chain in every training script of the reference
// (10_scripts/127_.../stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:134-136):
//     NormalizeIntensityd(nonzero=True, channel_wise=True) -> RandScaleIntensityd(0.1) -> RandShiftIntensityd(0.1)
// The three are MONAI 0.5 transforms (monai/transforms/intensity/array.py: NormalizeIntensity._normalize,
// ScaleIntensity.__call__, ShiftIntensity.__call__; MONAI is not part of /root/reference, oracle/monai_intensity.py
// restates them).  Per channel c:  m = x != 0;  mu = mean(x[m]);  sigma = std(x[m] - mu) (population; 1 if 0);
//     y = m ? ((x - mu) / sigma) * (1 + factor) + offset : offset          (an all-zero channel stays zero + offset)
// i.e. one masked reduction (count, sum, sum of squares) and one affine map  y = m ? a x + b : t  with
//     a = (1 + factor) / sigma,  b = offset - mu a,  t = offset.
// k_nz_stats reads the volume once (4 B/voxel), deterministic two-level reduction in double; k_affine_nz is the map on
// its own (8 B/voxel); the band-limited chain can apply the map while it loads x instead (mvtb_kspace_chain_pre_f32),
// so that chain + prologue move the chain's 8 B/voxel plus the 4 B/voxel of the statistics pass.
#include <math.h>

#include "mvtb_common.cuh"

namespace mvtb {

static const int kStatThreads = 256;

// grid (bx, channels): partial[(c * gridDim.x + bx) * 3 + {count, sum, sumsq}]
__global__ void __launch_bounds__(kStatThreads)
k_nz_stats(const float* __restrict__ in, size_t n_per_channel, double* __restrict__ partial) {
    const float* x = in + (size_t)blockIdx.y * n_per_channel;
    double cnt = 0.0, sm = 0.0, sq = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)x) & 15) == 0) {
        const size_t n4 = n_per_channel / 4;
        const float4* x4 = (const float4*)x;
        for (size_t g = i0; g < n4; g += stride) {
            const float4 v = x4[g];
            // a group of 4 in float (exact count; sums of 4 terms), accumulated in double
            const float c = (v.x != 0.f ? 1.f : 0.f) + (v.y != 0.f ? 1.f : 0.f) + (v.z != 0.f ? 1.f : 0.f) + (v.w != 0.f ? 1.f : 0.f);
            cnt += (double)c;
            sm += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
            sq += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
        }
        for (size_t e = n4 * 4 + i0; e < n_per_channel; e += stride) {
            const float v = x[e];
            cnt += v != 0.f ? 1.0 : 0.0; sm += (double)v; sq += (double)v * v;
        }
    } else {
        for (size_t e = i0; e < n_per_channel; e += stride) {
            const float v = x[e];
            cnt += v != 0.f ? 1.0 : 0.0; sm += (double)v; sq += (double)v * v;
        }
    }
    __shared__ double s_red[3][kStatThreads / 32];
    MVTB_UNROLL
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][w] = cnt; s_red[1][w] = sm; s_red[2][w] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int i = 0; i < kStatThreads / 32; ++i) { a += s_red[0][i]; b += s_red[1][i]; c += s_red[2][i]; }
        double* p = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3;
        p[0] = a; p[1] = b; p[2] = c;
    }
}

// one CTA per channel: partials in fixed order -> stats[c] = (count, mean, std); then (a, b, t) for the affine map
__global__ void __launch_bounds__(32)
k_nz_finish(const double* __restrict__ partial, int nblocks, double* __restrict__ stats, const float* __restrict__ scale,
            const float* __restrict__ shift, float* __restrict__ abt) {
    const int c = blockIdx.x, lane = threadIdx.x;
    double cnt = 0.0, sm = 0.0, sq = 0.0;
    for (int i = lane; i < nblocks; i += 32) {
        const double* p = partial + ((size_t)c * nblocks + i) * 3;
        cnt += p[0]; sm += p[1]; sq += p[2];
    }
    MVTB_UNROLL
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
        double mean = 0.0, sd = 1.0;
        if (cnt > 0.0) {
            mean = sm / cnt;
            double var = sq / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            sd = sqrt(var);
            if (sd == 0.0) sd = 1.0;                     // MONAI: a constant channel is only shifted
        }
        if (stats) { stats[3 * c] = cnt; stats[3 * c + 1] = mean; stats[3 * c + 2] = sd; }
        if (abt) {
            const double s = scale ? (double)scale[c] : 1.0, t = shift ? (double)shift[c] : 0.0;
            const double a = s / sd;
            abt[3 * c] = (float)a;
            abt[3 * c + 1] = (float)(t - mean * a);
            abt[3 * c + 2] = (float)t;
        }
    }
}

// y = x != 0 ? a x + b : t per channel; grid (bx, channels)
__global__ void __launch_bounds__(256)
k_affine_nz(const float* __restrict__ in, float* __restrict__ out, size_t n_per_channel, const float* __restrict__ abt) {
    const size_t base = (size_t)blockIdx.y * n_per_channel;
    const float a = __ldg(abt + 3 * blockIdx.y), b = __ldg(abt + 3 * blockIdx.y + 1), t = __ldg(abt + 3 * blockIdx.y + 2);
    const float* x = in + base;
    float* y = out + base;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0) {
        const size_t n4 = n_per_channel / 4;
        for (size_t g = i0; g < n4; g += stride) {
            const float4 v = ((const float4*)x)[g];
            float4 r;
            r.x = v.x != 0.f ? fmaf(a, v.x, b) : t;
            r.y = v.y != 0.f ? fmaf(a, v.y, b) : t;
            r.z = v.z != 0.f ? fmaf(a, v.z, b) : t;
            r.w = v.w != 0.f ? fmaf(a, v.w, b) : t;
            ((float4*)y)[g] = r;
        }
        for (size_t e = n4 * 4 + i0; e < n_per_channel; e += stride) y[e] = x[e] != 0.f ? fmaf(a, x[e], b) : t;
    } else {
        for (size_t e = i0; e < n_per_channel; e += stride) y[e] = x[e] != 0.f ? fmaf(a, x[e], b) : t;
    }
}

}  // namespace mvtb

using namespace mvtb;

static const int kStatBlocksCap = 148 * 8;

extern "C" size_t mvtb_intensity_scratch_bytes(int n_channels) {
    return n_channels > 0 ? sizeof(double) * 3 * (size_t)n_channels * (size_t)kStatBlocksCap : 0;
}

extern "C" int mvtb_intensity_prologue_coeffs_f32(const float* in, size_t n_per_channel, int n_channels, const float* scale,
                                                  const float* shift, double* stats_out, float* abt_out, void* scratch,
                                                  void* stream) {
    if (!in || !scratch || (!stats_out && !abt_out)) { set_error("intensity_prologue_coeffs: null argument"); return MVTB_EINVAL; }
    if (n_channels < 0 || n_channels > 65535) { set_error("intensity_prologue_coeffs: n_channels=%d", n_channels); return MVTB_EINVAL; }
    if (n_channels == 0) return MVTB_OK;
    if (n_per_channel == 0) { set_error("intensity_prologue_coeffs: empty channel"); return MVTB_EINVAL; }
    int per = kStatBlocksCap / n_channels;
    if (per < 1) per = 1;
    size_t want = (n_per_channel / 4 + kStatThreads - 1) / kStatThreads;
    if (want < 1) want = 1;
    const unsigned bx = (unsigned)(want < (size_t)per ? want : (size_t)per);
    MVTB_LAUNCH(k_nz_stats, dim3(bx, (unsigned)n_channels), dim3(kStatThreads), 0, stream, in, n_per_channel, (double*)scratch);
    MVTB_LAUNCH(k_nz_finish, dim3((unsigned)n_channels), dim3(32), 0, stream, (const double*)scratch, (int)bx, stats_out, scale, shift, abt_out);
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}

extern "C" int mvtb_intensity_affine_f32(const float* in, float* out, size_t n_per_channel, int n_channels, const float* abt,
                                         void* stream) {
    if (!in || !out || !abt) { set_error("intensity_affine: null argument"); return MVTB_EINVAL; }
    if (n_channels < 0 || n_channels > 65535) { set_error("intensity_affine: n_channels=%d", n_channels); return MVTB_EINVAL; }
    if (n_channels == 0 || n_per_channel == 0) return MVTB_OK;
    int per = kStatBlocksCap / n_channels;
    if (per < 1) per = 1;
    size_t want = (n_per_channel / 4 + 255) / 256;
    if (want < 1) want = 1;
    const unsigned bx = (unsigned)(want < (size_t)per ? want : (size_t)per);
    MVTB_LAUNCH(k_affine_nz, dim3(bx, (unsigned)n_channels), dim3(256), 0, stream, in, out, n_per_channel, abt);
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}
