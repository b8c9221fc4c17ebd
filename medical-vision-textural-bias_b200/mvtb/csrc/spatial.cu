// spatial.cu — the two index transforms between the resampled volume and the intensity prologue in every training script
// of the reference (10_scripts/127_.../stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:130-133):
//     RandSpatialCropd(roi_size=[128, 128, 64], random_size=False)  ->  RandFlipd(prob=0.5, spatial_axis=0)
// (validation: CenterSpatialCropd).  Both are MONAI 0.5 transforms (monai/transforms/croppad/array.py: SpatialCrop,
// CenterSpatialCrop, RandSpatialCrop; spatial/array.py: Flip = np.flip per channel; MONAI is not part of /root/reference,
// oracle/monai_spatial.py restates them).  Crop then flip is one gather:
//     out[c][i][j][k] = in[c][o0 + (f0 ? s0-1-i : i)][o1 + (f1 ? s1-1-j : j)][o2 + (f2 ? s2-1-k : k)]
// 8 B/voxel of the crop, nothing else; the host draws the offsets and the flip in MONAI's order (mvtb/spatial.py).
#include "mvtb_common.cuh"

namespace mvtb {

struct CropGeom {
    int iw, id;            // input W, D (row pitch and plane pitch in elements come from these)
    long long in_chan;     // elements per input channel
    int s0, s1, s2;        // output shape
    int o0, o1, o2;        // first input index kept on each axis
    int flip;              // bit a: spatial axis a is reversed
};

__global__ void __launch_bounds__(256)
k_crop_flip(const float* __restrict__ in, float* __restrict__ out, CropGeom g, long long n_rows_total) {
    // one warp per output row (c, i, j): the row's s2 elements are contiguous on both sides (reversed when axis 2 flips)
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = wid; r < n_rows_total; r += nw) {
        const long long c = r / ((long long)g.s0 * g.s1);
        const int rem = (int)(r - c * (long long)g.s0 * g.s1);
        const int i = rem / g.s1, j = rem - i * g.s1;
        const int si = g.o0 + ((g.flip & 1) ? g.s0 - 1 - i : i);
        const int sj = g.o1 + ((g.flip & 2) ? g.s1 - 1 - j : j);
        const float* src = in + c * g.in_chan + ((long long)si * g.iw + sj) * g.id + g.o2;
        float* dst = out + r * g.s2;
        if (g.flip & 4) {
            for (int k = lane; k < g.s2; k += 32) dst[k] = src[g.s2 - 1 - k];
        } else {
            for (int k = lane; k < g.s2; k += 32) dst[k] = src[k];
        }
    }
}

// the same gather for a batch of samples of one shape, each with its own window and flips (device arrays: the training
// loader draws them per sample); sample b reads in + b * in_sample, writes out + b * (C s0 s1 s2)
__global__ void __launch_bounds__(256)
k_crop_flip_batch(const float* __restrict__ in, float* __restrict__ out, CropGeom g, long long rows_per_sample, long long in_sample,
                  const int* __restrict__ offsets, const int* __restrict__ flips, long long n_rows_total) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long rr = wid; rr < n_rows_total; rr += nw) {
        const long long b = rr / rows_per_sample, r = rr - b * rows_per_sample;
        const int o0 = __ldg(offsets + 3 * b), o1 = __ldg(offsets + 3 * b + 1), o2 = __ldg(offsets + 3 * b + 2), fl = __ldg(flips + b);
        const long long c = r / ((long long)g.s0 * g.s1);
        const int rem = (int)(r - c * (long long)g.s0 * g.s1);
        const int i = rem / g.s1, j = rem - i * g.s1;
        const int si = o0 + ((fl & 1) ? g.s0 - 1 - i : i);
        const int sj = o1 + ((fl & 2) ? g.s1 - 1 - j : j);
        const float* src = in + b * in_sample + c * g.in_chan + ((long long)si * g.iw + sj) * g.id + o2;
        float* dst = out + rr * g.s2;
        if (fl & 4) {
            for (int k = lane; k < g.s2; k += 32) dst[k] = src[g.s2 - 1 - k];
        } else {
            for (int k = lane; k < g.s2; k += 32) dst[k] = src[k];
        }
    }
}

}  // namespace mvtb

using namespace mvtb;

extern "C" int mvtb_crop_flip_f32(const float* in, float* out, int n_channels, const int32_t* in_shape, const int32_t* out_shape,
                                  const int32_t* offset, int flip_axes_mask, void* stream) {
    if (!in || !out || !in_shape || !out_shape || !offset) { set_error("crop_flip: null argument"); return MVTB_EINVAL; }
    if (in == out) { set_error("crop_flip: in-place is not supported"); return MVTB_EINVAL; }
    if (n_channels < 0 || (flip_axes_mask & ~7)) { set_error("crop_flip: bad channel count or flip mask"); return MVTB_EINVAL; }
    for (int a = 0; a < 3; ++a) {
        if (in_shape[a] < 1 || out_shape[a] < 0 || offset[a] < 0 || (long long)offset[a] + out_shape[a] > in_shape[a]) {
            set_error("crop_flip: axis %d: offset %d + size %d does not fit in %d", a, offset[a], out_shape[a], in_shape[a]);
            return MVTB_EINVAL;
        }
    }
    const long long rows = (long long)n_channels * out_shape[0] * out_shape[1];
    if (rows == 0 || out_shape[2] == 0) return MVTB_OK;
    CropGeom g;
    g.iw = in_shape[1]; g.id = in_shape[2];
    g.in_chan = (long long)in_shape[0] * in_shape[1] * in_shape[2];
    g.s0 = out_shape[0]; g.s1 = out_shape[1]; g.s2 = out_shape[2];
    g.o0 = offset[0]; g.o1 = offset[1]; g.o2 = offset[2];
    g.flip = flip_axes_mask;
    long long blocks = (rows + 7) / 8;                      // 8 warps per CTA
    if (blocks > 148 * 16) blocks = 148 * 16;
    MVTB_LAUNCH(k_crop_flip, dim3((unsigned)blocks), dim3(256), 0, stream, in, out, g, rows);
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}

// offsets_dev[3 n_samples], flips_dev[n_samples]: int32 on the device; every window must fit (the caller draws them with
// MONAI's bounds, 0 <= o <= N - s); not checked here, the arrays live on the device.
extern "C" int mvtb_crop_flip_batch_f32(const float* in, float* out, int n_samples, int n_channels, const int32_t* in_shape,
                                        const int32_t* out_shape, const int32_t* offsets_dev, const int32_t* flips_dev, void* stream) {
    if (!in || !out || !in_shape || !out_shape || !offsets_dev || !flips_dev) { set_error("crop_flip_batch: null argument"); return MVTB_EINVAL; }
    if (in == out) { set_error("crop_flip_batch: in-place is not supported"); return MVTB_EINVAL; }
    if (n_samples < 0 || n_channels < 0) { set_error("crop_flip_batch: negative count"); return MVTB_EINVAL; }
    for (int a = 0; a < 3; ++a)
        if (in_shape[a] < 1 || out_shape[a] < 0 || out_shape[a] > in_shape[a]) { set_error("crop_flip_batch: axis %d: size %d of %d", a, out_shape[a], in_shape[a]); return MVTB_EINVAL; }
    const long long rows_per_sample = (long long)n_channels * out_shape[0] * out_shape[1];
    const long long rows = rows_per_sample * n_samples;
    if (rows == 0 || out_shape[2] == 0) return MVTB_OK;
    CropGeom g;
    g.iw = in_shape[1]; g.id = in_shape[2];
    g.in_chan = (long long)in_shape[0] * in_shape[1] * in_shape[2];
    g.s0 = out_shape[0]; g.s1 = out_shape[1]; g.s2 = out_shape[2];
    g.o0 = g.o1 = g.o2 = 0; g.flip = 0;
    long long blocks = (rows + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    MVTB_LAUNCH(k_crop_flip_batch, dim3((unsigned)blocks), dim3(256), 0, stream, in, out, g, rows_per_sample, g.in_chan * n_channels,
                offsets_dev, flips_dev, rows);
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}
