// pointwise.cuh — the k-space pointwise stage for one bin, shared by the band-limited path.
// (kspace_chain.cu keeps its own column-factored version of the same arithmetic.)
//
// With fftshift-ed index i_d per axis and its negation i'_d = (2 floor(N/2) - i_d) mod N:
//   M_eff = (M(i) + M(i')) / 2              real part after an asymmetric mask
//   spike at this bin      : + (new - k_old) / 2,      k_old = M_eff(f_s) K  (what the previous stage's REAL output
//                                                       holds at that bin: GibbsNoise -> KSpaceSpikeNoise in sequence)
//   spike at the conjugate : + conj(new - k_old) / 2,  k_old = M_eff(f_s) conj K   (K(f_s) = conj K(-f_s))
//   self-conjugate spike   : bin becomes Re(new)
//   new = amp * k_old / |k_old|  (amp if k_old == 0: angle(0) = 0, F:384/F:928)
//   wrap: x alpha for every odd i_d among the trailing wrap_naxes axes (F:509-511)
#pragma once
#include <math.h>

#include "mvtb_common.cuh"

namespace mvtb {

__device__ __forceinline__ long long pw_mask_term(int kind, int i, int n) {
    const long long d = kind == MVTB_MASK_DISK ? (long long)(i - n / 2) : (long long)(2 * i - (n - 1));
    return d * d;
}

__device__ __forceinline__ cf pw_spike_value(cf ko, float amp) {
    const float mag = hypotf(ko.x, ko.y);
    if (mag > 0.f) return cmk(amp * (ko.x / mag), amp * (ko.y / mag));
    return cmk(amp, 0.f);
}

// ish[a]: fftshift-ed index on FFT axis a (axis 0 = last axis), shape[a] its length.
__device__ __forceinline__ cf pointwise_bin(const DescDev& d, int ndim, const int* shape, const int* ish, cf K, float scale) {
    int ineg[MVTB_MAX_FFT_DIMS];
    long long qp = 0, qn = 0;
    float w = scale;
    for (int a = 0; a < ndim; ++a) {
        const int n = shape[a];
        ineg[a] = (2 * (n / 2) - ish[a] + n) % n;
        if (d.mask_kind != MVTB_MASK_NONE && a < d.mask_ndim) {
            qp += pw_mask_term(d.mask_kind, ish[a], n);
            qn += pw_mask_term(d.mask_kind, ineg[a], n);
        }
        if (a < d.wrap_naxes && (ish[a] & 1)) w *= d.wrap_alpha;
    }
    float meff = 1.f;
    if (d.mask_kind != MVTB_MASK_NONE) {
        const int kp = (qp <= d.thr ? 1 : 0) ^ d.inside_off;
        const int kn = (qn <= d.thr ? 1 : 0) ^ d.inside_off;
        meff = 0.5f * (float)(kp + kn);
    }
    cf acc = cscale(K, meff);
    for (int s = 0; s < d.n_spikes; ++s) {
        bool isp = true, isn = true;
        for (int a = 0; a < ndim; ++a) {
            isp = isp && ish[a] == d.sp[s].idx[a];
            isn = isn && ineg[a] == d.sp[s].idx[a];
        }
        if (!isp && !isn) continue;
        const float ms = d.sp[s].meff_at_spike;
        if (isp && isn) {
            const cf ko = cscale(K, ms);
            const cf nw = pw_spike_value(ko, d.sp[s].amp);
            acc.x += nw.x - ko.x;
            acc.y -= ko.y;
        } else if (isp) {
            const cf ko = cscale(K, ms);
            const cf nw = pw_spike_value(ko, d.sp[s].amp);
            acc.x += 0.5f * (nw.x - ko.x);
            acc.y += 0.5f * (nw.y - ko.y);
        } else {
            const cf ko = cscale(cconj(K), ms);
            const cf nw = pw_spike_value(ko, d.sp[s].amp);
            acc.x += 0.5f * (nw.x - ko.x);
            acc.y -= 0.5f * (nw.y - ko.y);
        }
    }
    return cscale(acc, w);
}

}  // namespace mvtb
