// dice.cu — the reductions that follow the hot path in every training / evaluation step of the reference:
//   DiceLoss(to_onehot_y=False, sigmoid=True, squared_pred=True)       10_scripts/127_.../...FLAIR.py:216
//   Activations(sigmoid) -> AsDiscrete(threshold 0.5) -> DiceMetric    ...FLAIR.py:266-283, utils.py:313-415
//   (twice more per step in Gibbs_GD's finite differences, 350_stylized_layers/gibbs0p7_layer_domain_GD.py:252-269)
// MONAI 0.5 (monai/losses/dice.py, monai/metrics/meandice.py; not part of /root/reference, restated in
// oracle/monai_losses.py) does this with ~10 elementwise / reduction passes over the (B, C, H, W, D) logits.  Here ONE
// pass over logits x and target t produces, per (b, c) volume, the six sums both need:
//   [0] sum t p   [1] sum p^2   [2] sum t^2        (loss;  p = sigmoid(x), or x itself when from_logits == 0)
//   [3] sum t q   [4] sum q     [5] sum t           (metric; q = [p >= 0.5] = [x >= 0])
// in double with a fixed-order two-level reduction (deterministic), and ONE pass produces the loss gradient
//   dL/dx_i = (A t_i + B p_i) p_i (1 - p_i),   A, B per volume from the sums (host / torch side, a few scalars).
#include <math.h>

#include "mvtb_common.cuh"

namespace mvtb {

static const int kDiceThreads = 256;

__device__ __forceinline__ float dice_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }   // torch.sigmoid's formula in fp32

__global__ void __launch_bounds__(kDiceThreads)
k_dice_sums(const float* __restrict__ x, const float* __restrict__ t, size_t n_per_vol, int from_logits, double* __restrict__ partial) {
    const float* xv = x + (size_t)blockIdx.y * n_per_vol;
    const float* tv = t + (size_t)blockIdx.y * n_per_vol;
    double s[6] = {0, 0, 0, 0, 0, 0};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_per_vol; e += stride) {
        const float xi = xv[e], ti = tv[e];
        const float p = from_logits ? dice_sigmoid(xi) : xi;
        const float q = p >= 0.5f ? 1.f : 0.f;
        s[0] += (double)(ti * p);
        s[1] += (double)(p * p);
        s[2] += (double)(ti * ti);
        s[3] += (double)(ti * q);
        s[4] += (double)q;
        s[5] += (double)ti;
    }
    __shared__ double red[6][kDiceThreads / 32];
    MVTB_UNROLL
    for (int k = 0; k < 6; ++k) {
        MVTB_UNROLL
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 6; ++k) red[k][w] = s[k];
    __syncthreads();
    if (threadIdx.x < 6) {
        double a = 0.0;
        for (int i = 0; i < kDiceThreads / 32; ++i) a += red[threadIdx.x][i];
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 6 + threadIdx.x] = a;
    }
}

__global__ void __launch_bounds__(32)
k_dice_finish(const double* __restrict__ partial, int nblocks, double* __restrict__ sums) {
    const int v = blockIdx.x, lane = threadIdx.x;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = lane; i < nblocks; i += 32)
        for (int k = 0; k < 6; ++k) s[k] += partial[((size_t)v * nblocks + i) * 6 + k];
    MVTB_UNROLL
    for (int k = 0; k < 6; ++k) {
        MVTB_UNROLL
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
    if (lane == 0)
        for (int k = 0; k < 6; ++k) sums[6 * v + k] = s[k];
}

// grad_i = (A t_i + B p_i) * (from_logits ? p_i (1 - p_i) : 1),  coef[2 v] = A, coef[2 v + 1] = B
__global__ void __launch_bounds__(256)
k_dice_grad(const float* __restrict__ x, const float* __restrict__ t, size_t n_per_vol, int from_logits,
            const float* __restrict__ coef, float* __restrict__ g) {
    const size_t base = (size_t)blockIdx.y * n_per_vol;
    const float A = __ldg(coef + 2 * blockIdx.y), B = __ldg(coef + 2 * blockIdx.y + 1);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_per_vol; e += stride) {
        const float xi = x[base + e], ti = t[base + e];
        const float p = from_logits ? dice_sigmoid(xi) : xi;
        const float d = fmaf(A, ti, B * p);
        g[base + e] = from_logits ? d * (p * (1.f - p)) : d;
    }
}

// sum (a - b)^2 in double, one partial per CTA (fixed order): the frequency-consistency loss of the reconstruction GAN
// (50_reconstruction/reconGan/reconGan_freq.py:134-140) is, by Parseval, H W times the image-domain mean squared error
__global__ void __launch_bounds__(kDiceThreads)
k_sqdiff(const float* __restrict__ a, const float* __restrict__ b, size_t n, double* __restrict__ partial) {
    double s = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const float d = a[e] - b[e];
        s += (double)d * (double)d;
    }
    __shared__ double red[kDiceThreads / 32];
    MVTB_UNROLL
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < kDiceThreads / 32; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(32)
k_sum_partials(const double* __restrict__ partial, int n, double* __restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
    MVTB_UNROLL
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = s;
}

}  // namespace mvtb

using namespace mvtb;

static const int kDiceBlocksCap = 148 * 8;

extern "C" size_t mvtb_dice_scratch_bytes(int n_vols) { return n_vols > 0 ? sizeof(double) * 6 * (size_t)n_vols * (size_t)kDiceBlocksCap : 0; }

static unsigned dice_bx(size_t n_per_vol, int n_vols) {
    int per = kDiceBlocksCap / n_vols;
    if (per < 1) per = 1;
    size_t want = (n_per_vol / 4 + kDiceThreads - 1) / kDiceThreads;
    if (want < 1) want = 1;
    return (unsigned)(want < (size_t)per ? want : (size_t)per);
}

extern "C" int mvtb_dice_sums_f32(const float* x, const float* target, size_t n_per_vol, int n_vols, int from_logits,
                                  double* sums_out, void* scratch, void* stream) {
    if (!x || !target || !sums_out || !scratch) { set_error("dice_sums: null argument"); return MVTB_EINVAL; }
    if (n_vols < 0 || n_vols > 65535) { set_error("dice_sums: n_vols=%d", n_vols); return MVTB_EINVAL; }
    if (n_vols == 0) return MVTB_OK;
    if (n_per_vol == 0) { MVTB_CUDA(cudaMemsetAsync(sums_out, 0, sizeof(double) * 6 * (size_t)n_vols, (cudaStream_t)stream)); return MVTB_OK; }
    const unsigned bx = dice_bx(n_per_vol, n_vols);
    MVTB_LAUNCH(k_dice_sums, dim3(bx, (unsigned)n_vols), dim3(kDiceThreads), 0, stream, x, target, n_per_vol, from_logits, (double*)scratch);
    MVTB_LAUNCH(k_dice_finish, dim3((unsigned)n_vols), dim3(32), 0, stream, (const double*)scratch, (int)bx, sums_out);
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}

extern "C" int mvtb_dice_grad_f32(const float* x, const float* target, size_t n_per_vol, int n_vols, int from_logits,
                                  const float* coef, float* grad_out, void* stream) {
    if (!x || !target || !coef || !grad_out) { set_error("dice_grad: null argument"); return MVTB_EINVAL; }
    if (n_vols < 0 || n_vols > 65535) { set_error("dice_grad: n_vols=%d", n_vols); return MVTB_EINVAL; }
    if (n_vols == 0 || n_per_vol == 0) return MVTB_OK;
    const unsigned bx = dice_bx(n_per_vol, n_vols);
    MVTB_LAUNCH(k_dice_grad, dim3(bx, (unsigned)n_vols), dim3(256), 0, stream, x, target, n_per_vol, from_logits, coef, grad_out);
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}

extern "C" int mvtb_sqdiff_sum_f32(const float* a, const float* b, size_t n, double* sum_out, void* scratch, void* stream) {
    if (!a || !b || !sum_out || !scratch) { set_error("sqdiff_sum: null argument"); return MVTB_EINVAL; }
    if (n == 0) { MVTB_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(double), (cudaStream_t)stream)); return MVTB_OK; }
    size_t want = (n / 4 + kDiceThreads - 1) / kDiceThreads;
    if (want < 1) want = 1;
    const unsigned bx = (unsigned)(want < (size_t)kDiceBlocksCap ? want : (size_t)kDiceBlocksCap);
    MVTB_LAUNCH(k_sqdiff, dim3(bx), dim3(kDiceThreads), 0, stream, a, b, n, (double*)scratch);
    MVTB_LAUNCH(k_sum_partials, dim3(1), dim3(32), 0, stream, (const double*)scratch, (int)bx, sum_out);
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}
