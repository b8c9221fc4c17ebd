// spike_fast.cu — k-space spikes without an FFT.
//
// RandPlaneWaves_ellipsoid (F:370-393), KSpaceSpikeNoise (F:906-945) and spike_layer (S:143-151) rewrite
// log|k| at a few k-space voxels and return the real part of the inverse transform.  Only those bins (and
// their conjugate partners) change, so with k_old = K(f_s) = sum_n x[n] exp(-2 pi i f_s.n/N)
//     out = x + Re( (amp k_old/|k_old| - k_old) exp(+2 pi i f_s.n/N) ) / N          
// i.e. one DFT coefficient per spike (a reduction over the volume) and one plane-wave axpy:
//   k_spike_reduce   reads the volume once, per-CTA partial sums (fixed order: deterministic)
//   k_spike_apply    sums the partials, forms the plane-wave amplitude, out = x + plane waves
// 12 B/voxel of traffic at most (8 when the second read hits L2) against ~40 B/voxel and 5 launches for the
// general FFT path.  Taken when a descriptor has spikes only (no mask, no wrap); any FFT rank 2..4.
#include <math.h>
#include <string.h>

#include <vector>

#include "mvtb_common.cuh"

namespace mvtb {

int plan_stage_upload(mvtb_plan* p, const void* src, size_t bytes, void* stream, void** dptr);   // plan.cu

static const int kSpThreads = 256;
static const int kSpGroupRows = 64;        // rows per pass over the CTA's span (8 warps x kSpUnroll rows)
static const int kSpRowsPerCta = 256;      // rows of the last axis per CTA: four passes share one table load and setup
                                           // (64 rows per CTA = 40 KB of volume per 5 KB table load: 15.6 us reduce, 19.2 us apply)

struct SpVol {
    int n;
    int f[MVTB_MAX_SPIKES][MVTB_MAX_FFT_DIMS];     // signed frequency per FFT axis, axis 0 = last
    float amp[MVTB_MAX_SPIKES];
};

struct SpGeom {
    int ndim;
    int shape[MVTB_MAX_FFT_DIMS];                  // axis 0 = last
    long long rows;                                // product of shape[1..]
    float scale;                                   // 1 / prod(shape)
    int toff[MVTB_MAX_FFT_DIMS];                   // offset of each axis' phase table inside a spike's table block
    int tlen;                                      // sum of the axis lengths
    int nmax;                                      // largest spike count of any volume in the call
};

__device__ __forceinline__ cf sp_unit(int f, int n, int N, float sign) {
    int m = (int)(((long long)f * n) % N);
    if (m < 0) m += N;
    float s, c;
    sincospif(2.0f * (float)m / (float)N, &s, &c);
    return cmk(c, sign * s);
}

// exp(sign * 2 pi i f_a n / N_a) for every axis and spike into shared memory:
// st[s * g.tlen + g.toff[a] + n], n < shape[a]
__device__ __forceinline__ void sp_tables(cf* st, const SpVol& sv, const SpGeom& g, float sign, int tid, int nthr) {
    for (int e = tid; e < sv.n * g.tlen; e += nthr) {
        const int s = e / g.tlen, r = e - s * g.tlen;
        int a = 0;
        while (a + 1 < g.ndim && r >= g.toff[a + 1]) ++a;
        st[e] = sp_unit(sv.f[s][a], r - g.toff[a], g.shape[a], sign);
    }
}

// one CTA per volume: the phase tables exp(-2 pi i f_a n / N_a) of every spike and axis -> global memory
__global__ void __launch_bounds__(256)
k_spike_tables(cf* __restrict__ tabs, SpGeom g, const SpVol* __restrict__ vols, int shared_desc) {
    const int vol = blockIdx.x;
    const SpVol& sv = vols[shared_desc ? 0 : vol];
    sp_tables(tabs + (size_t)vol * MVTB_MAX_SPIKES * g.tlen, sv, g, -1.f, threadIdx.x, blockDim.x);
}

// copy a volume's tables into shared memory (conjugated for the inverse direction)
__device__ __forceinline__ void sp_load_tables(cf* st, const cf* __restrict__ tabs, int n, const SpGeom& g, bool conj,
                                               int tid, int nthr) {
    for (int e = tid; e < n * g.tlen; e += nthr) {
        const cf t = __ldg(tabs + e);
        st[e] = conj ? cmk(t.x, -t.y) : t;
    }
}

// product over the axes >= 1 of the table entries for flattened row index `row` (32-bit: rows < 2^31)
__device__ __forceinline__ cf sp_row_phase(const cf* st_s, const SpGeom& g, unsigned row) {
    cf e = cmk(1.f, 0.f);
    for (int a = 1; a < g.ndim; ++a) {
        const unsigned q = row / (unsigned)g.shape[a];
        const unsigned na = row - q * (unsigned)g.shape[a];
        row = q;
        e = cmul(e, st_s[g.toff[a] + na]);
    }
    return e;
}

// Per CTA: kSpRowsPerCta consecutive rows = one contiguous span of the volume, walked as a flat array with
// kSpUnroll independent coalesced loads in flight per thread; the phase of each row's outer axes is
// precomputed once per CTA in shared memory.
static const int kSpUnroll = 8;            // 16 rows per warp and step is slower (26.7 vs 18.9 us per volume in k_spike_apply)

__device__ __forceinline__ void sp_row_phases(cf* srow, const cf* st, const SpVol& sv, const SpGeom& g,
                                              long long row0, int nrows, int tid, int nthr) {
    for (int e = tid; e < sv.n * kSpRowsPerCta; e += nthr) {
        const int s = e / kSpRowsPerCta, r = e - s * kSpRowsPerCta;
        srow[e] = r < nrows ? sp_row_phase(st + s * g.tlen, g, (unsigned)(row0 + r)) : cmk(0.f, 0.f);
    }
}

// grid = (ctas per volume, volumes); partial[(vol * gridDim.x + cta) * MVTB_MAX_SPIKES + s]
__global__ void __launch_bounds__(256)
k_spike_reduce(const float* __restrict__ x, cf* __restrict__ partial, SpGeom g, const SpVol* __restrict__ vols, int shared_desc,
               const cf* __restrict__ tabs) {
    MVTB_DYN_SMEM(smem_raw);
    cf* st = (cf*)smem_raw;                                    // [n][tlen] axis phase tables
    cf* srow = st + g.nmax * g.tlen;                           // [n][kSpRowsPerCta] outer-axes phase per row
    __shared__ cf s_red[8][MVTB_MAX_SPIKES];
    const int vol = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const SpVol& sv = vols[shared_desc ? 0 : vol];
    const int n0 = g.shape[0];
    const long long row0 = (long long)blockIdx.x * kSpRowsPerCta;
    const int nrows = (int)(g.rows - row0 < kSpRowsPerCta ? g.rows - row0 : kSpRowsPerCta);
    sp_load_tables(st, tabs + (size_t)vol * MVTB_MAX_SPIKES * g.tlen, sv.n, g, false, tid, blockDim.x);
    __syncthreads();
    sp_row_phases(srow, st, sv, g, row0, nrows, tid, blockDim.x);
    __syncthreads();
    cf acc[MVTB_MAX_SPIKES];
    MVTB_UNROLL
    for (int s = 0; s < MVTB_MAX_SPIKES; ++s) acc[s] = cmk(0.f, 0.f);
    // per pass, warp w owns kSpUnroll consecutive rows of a kSpGroupRows-row group: that many independent coalesced loads per lane and step
    const float* xb = x + ((size_t)vol * g.rows + row0) * n0;
    for (int grp = 0; grp < kSpRowsPerCta && grp < nrows; grp += kSpGroupRows) {
        const int rbase = grp + wp * kSpUnroll;
        const float* xr[kSpUnroll];                              // rows past the end alias the last row; their
        MVTB_UNROLL                                              // row phase is 0, so they contribute nothing
        for (int k = 0; k < kSpUnroll; ++k) xr[k] = xb + (size_t)(rbase + k < nrows ? rbase + k : nrows - 1) * n0;
        for (int s = 0; s < sv.n; ++s) {
            cf part[kSpUnroll];
            MVTB_UNROLL
            for (int k = 0; k < kSpUnroll; ++k) part[k] = cmk(0.f, 0.f);
            for (int i = lane; i < n0; i += 32) {
                float v[kSpUnroll];
                MVTB_UNROLL
                for (int k = 0; k < kSpUnroll; ++k) v[k] = xr[k][i];    // kSpUnroll unconditional loads in flight
                const cf t = st[s * g.tlen + i];
                MVTB_UNROLL
                for (int k = 0; k < kSpUnroll; ++k) { part[k].x = fmaf(v[k], t.x, part[k].x); part[k].y = fmaf(v[k], t.y, part[k].y); }
            }
            cf a_ = cmk(0.f, 0.f);
            MVTB_UNROLL
            for (int k = 0; k < kSpUnroll; ++k) a_ = cadd(a_, cmul(part[k], srow[s * kSpRowsPerCta + rbase + k]));
            MVTB_UNROLL
            for (int s2 = 0; s2 < MVTB_MAX_SPIKES; ++s2) if (s2 == s) acc[s2] = cadd(acc[s2], a_);
        }
    }
    MVTB_UNROLL
    for (int s = 0; s < MVTB_MAX_SPIKES; ++s) {
        MVTB_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            acc[s].x += __shfl_xor_sync(0xffffffffu, acc[s].x, o);
            acc[s].y += __shfl_xor_sync(0xffffffffu, acc[s].y, o);
        }
        if (lane == 0) s_red[wp][s] = acc[s];
    }
    __syncthreads();
    if (tid < MVTB_MAX_SPIKES) {
        cf t = cmk(0.f, 0.f);
        for (int w = 0; w < 8; ++w) t = cadd(t, s_red[w][tid]);
        partial[((size_t)vol * gridDim.x + blockIdx.x) * MVTB_MAX_SPIKES + tid] = t;
    }
}

__device__ __forceinline__ void sp_atomic_min(float* addr, float v) {
    v += 0.0f;
    if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v));
    else atomicMax((unsigned*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void sp_atomic_max(float* addr, float v) {
    v += 0.0f;
    if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
    else atomicMin((unsigned*)addr, __float_as_uint(v));
}

// one CTA per volume: K(f_s) = sum of the per-CTA partials in a fixed order (double accumulation), then the
// plane-wave amplitude Delta_s / N = (amp K/|K| - K) / N  (amp if K == 0: angle(0) = 0)
__global__ void __launch_bounds__(256)
k_spike_finalize(const cf* __restrict__ partial, int n_partial, cf* __restrict__ delta, SpGeom g,
                 const SpVol* __restrict__ vols, int shared_desc) {
    __shared__ double s_sum[kSpThreads][2];
    const int vol = blockIdx.x, tid = threadIdx.x;
    const SpVol& sv = vols[shared_desc ? 0 : vol];
    for (int s = 0; s < sv.n; ++s) {
        double ax = 0.0, ay = 0.0;
        for (int i = tid; i < n_partial; i += blockDim.x) {
            const cf v = partial[((size_t)vol * n_partial + i) * MVTB_MAX_SPIKES + s];
            ax += (double)v.x;
            ay += (double)v.y;
        }
        s_sum[tid][0] = ax;
        s_sum[tid][1] = ay;
        __syncthreads();
        for (int o = kSpThreads / 2; o > 0; o >>= 1) {
            if (tid < o) { s_sum[tid][0] += s_sum[tid + o][0]; s_sum[tid][1] += s_sum[tid + o][1]; }
            __syncthreads();
        }
        if (tid == 0) {
            const cf k = cmk((float)s_sum[0][0], (float)s_sum[0][1]);
            const float mag = hypotf(k.x, k.y);
            const cf nw = mag > 0.f ? cmk(sv.amp[s] * (k.x / mag), sv.amp[s] * (k.y / mag)) : cmk(sv.amp[s], 0.f);
            delta[(size_t)vol * MVTB_MAX_SPIKES + s] = cmk((nw.x - k.x) * g.scale, (nw.y - k.y) * g.scale);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_spike_apply(const float* __restrict__ x, float* __restrict__ out, const cf* __restrict__ delta,
              SpGeom g, const SpVol* __restrict__ vols, int shared_desc, const cf* __restrict__ tabs,
              float* __restrict__ minmax, int vols_per_sample, int vol_base) {
    MVTB_DYN_SMEM(smem_raw);
    cf* st = (cf*)smem_raw;                                    // [n][tlen] axis phase tables
    cf* srow = st + g.nmax * g.tlen;                           // [n][kSpRowsPerCta]
    __shared__ cf s_delta[MVTB_MAX_SPIKES];
    const int vol = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const SpVol& sv = vols[shared_desc ? 0 : vol];
    sp_load_tables(st, tabs + (size_t)vol * MVTB_MAX_SPIKES * g.tlen, sv.n, g, true, tid, blockDim.x);
    if (tid < MVTB_MAX_SPIKES) s_delta[tid] = tid < sv.n ? delta[(size_t)vol * MVTB_MAX_SPIKES + tid] : cmk(0.f, 0.f);
    __syncthreads();
    const int n0 = g.shape[0];
    const long long row0 = (long long)blockIdx.x * kSpRowsPerCta;
    const int nrows = (int)(g.rows - row0 < kSpRowsPerCta ? g.rows - row0 : kSpRowsPerCta);
    sp_row_phases(srow, st, sv, g, row0, nrows, tid, blockDim.x);
    __syncthreads();
    for (int e = tid; e < sv.n * kSpRowsPerCta; e += blockDim.x)      // fold Delta_s into the row phase
        srow[e] = cmul(srow[e], s_delta[e / kSpRowsPerCta]);
    __syncthreads();
    const float* xb = x + ((size_t)vol * g.rows + row0) * n0;
    float* ob = out + ((size_t)vol * g.rows + row0) * n0;
    float lo = __int_as_float(0x7f800000), hi = __int_as_float((int)0xff800000u);
    for (int grp = 0; grp < kSpRowsPerCta && grp < nrows; grp += kSpGroupRows) {
        const int rbase = grp + wp * kSpUnroll;
        const float* xr[kSpUnroll];
        MVTB_UNROLL
        for (int k = 0; k < kSpUnroll; ++k) xr[k] = xb + (size_t)(rbase + k < nrows ? rbase + k : nrows - 1) * n0;
        float v[kSpUnroll], vn[kSpUnroll];
        MVTB_UNROLL
        for (int k = 0; k < kSpUnroll; ++k) v[k] = lane < n0 ? xr[k][lane] : 0.f;
        for (int i = lane; i < n0; i += 32) {
            // software pipeline: the loads of the next 32 columns are in flight while this step computes and stores
            const int inext = i + 32;
            MVTB_UNROLL
            for (int k = 0; k < kSpUnroll; ++k) vn[k] = inext < n0 ? xr[k][inext] : 0.f;
            for (int s = 0; s < sv.n; ++s) {
                const cf t = st[s * g.tlen + i];
                MVTB_UNROLL
                for (int k = 0; k < kSpUnroll; ++k) {
                    const cf de = srow[s * kSpRowsPerCta + rbase + k];  // Delta_s * row phase (broadcast)
                    v[k] += de.x * t.x - de.y * t.y;
                }
            }
            MVTB_UNROLL
            for (int k = 0; k < kSpUnroll; ++k) {
                if (rbase + k < nrows) {
                    ob[(size_t)(rbase + k) * n0 + i] = v[k];
                    lo = fminf(lo, v[k]);
                    hi = fmaxf(hi, v[k]);
                }
            }
            MVTB_UNROLL
            for (int k = 0; k < kSpUnroll; ++k) v[k] = vn[k];
        }
    }
    if (minmax != nullptr) {
        __shared__ float s_lo[8], s_hi[8];
        MVTB_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { s_lo[wp] = lo; s_hi[wp] = hi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
            float* mm = minmax + 2 * ((vol_base + vol) / vols_per_sample);
            sp_atomic_min(mm, lo);
            sp_atomic_max(mm + 1, hi);
        }
    }
}

// spikes only: no mask, no wrap, at least one spike in every descriptor
bool spike_fast_eligible(const mvtb_plan* p, const mvtb_chain_desc* desc, int n_desc) {
    if (p->opt_path == MVTB_PATH_GENERAL) return false;
    long long tlen = 0, rows = 1;
    for (int a = 0; a < p->ndim; ++a) { tlen += p->shape[a]; if (a >= 1) rows *= p->shape[a]; }
    if (sizeof(cf) * MVTB_MAX_SPIKES * (size_t)(tlen + kSpRowsPerCta) > 200 * 1024 || rows > 0x7fffffffLL) return false;   // tables must fit shared memory
    int total = 0;
    for (int i = 0; i < n_desc; ++i) {
        const mvtb_chain_desc& d = desc[i];
        if (d.mask_kind != MVTB_MASK_NONE || d.wrap_naxes != 0) return false;
        if (d.n_spikes < 0 || d.n_spikes > MVTB_MAX_SPIKES) return false;     // volumes without a spike are copied
        total += d.n_spikes;
        for (int s = 0; s < d.n_spikes; ++s)
            for (int a = 0; a < p->ndim; ++a) {
                const int idx = d.spikes[s].idx[a];
                if (idx < 0 || idx >= p->shape[p->ndim - 1 - a]) return false;      // the general path reports it
                for (int s2 = 0; s2 < s; ++s2) {
                    bool same = true;
                    for (int b = 0; b < p->ndim; ++b) same = same && d.spikes[s2].idx[b] == d.spikes[s].idx[b];
                    if (same) return false;
                }
            }
    }
    if (total == 0) return false;
    return true;
}

int spike_fast_chain(mvtb_plan* p, const float* in, float* out, int n_volumes, const mvtb_chain_desc* desc, int n_desc,
                     float* minmax_out, int vols_per_sample, void* stream) {
    SpGeom g;
    memset(&g, 0, sizeof(g));
    g.ndim = p->ndim;
    g.rows = 1;
    double tot = 1.0;
    for (int a = 0; a < p->ndim; ++a) {
        g.shape[a] = p->shape[a];
        tot *= (double)p->shape[a];
        if (a >= 1) g.rows *= p->shape[a];
    }
    g.scale = (float)(1.0 / tot);
    g.tlen = 0;
    for (int a = 0; a < p->ndim; ++a) { g.toff[a] = g.tlen; g.tlen += p->shape[a]; }
    const long long ctas_ll = (g.rows + kSpRowsPerCta - 1) / kSpRowsPerCta;
    if (ctas_ll > 0x7fffffffLL || n_volumes > 65535) { set_error("spike fast path: grid too large"); return MVTB_EUNSUPPORTED; }
    const int ctas = (int)ctas_ll;

    std::vector<SpVol> hv((size_t)n_desc);
    for (int i = 0; i < n_desc; ++i) {
        SpVol& v = hv[i];
        memset(&v, 0, sizeof(v));
        v.n = desc[i].n_spikes;
        for (int s = 0; s < v.n; ++s) {
            for (int a = 0; a < p->ndim; ++a) {
                const int n = p->shape[a];
                v.f[s][a] = desc[i].spikes[s].idx[p->ndim - 1 - a] - n / 2;     // user order: outermost first
            }
            v.amp[s] = desc[i].spikes[s].amplitude;
        }
    }
    g.nmax = 1;
    for (int i = 0; i < n_desc; ++i) g.nmax = hv[i].n > g.nmax ? hv[i].n : g.nmax;
    void* dvp = nullptr;
    int rc = plan_stage_upload(p, hv.data(), hv.size() * sizeof(SpVol), stream, &dvp);
    if (rc != MVTB_OK) return rc;
    const SpVol* dv = (const SpVol*)dvp;
    const int shared_desc = n_desc == 1 ? 1 : 0;

    // workspace per volume: per-CTA partial sums, the Delta_s, and the phase tables (all in the plan's workspace)
    const size_t per_partial = sizeof(cf) * MVTB_MAX_SPIKES * (size_t)ctas;
    const size_t per_delta = sizeof(cf) * MVTB_MAX_SPIKES;
    const size_t per_tab = sizeof(cf) * MVTB_MAX_SPIKES * (size_t)g.tlen;
    const size_t need = per_partial + per_delta + per_tab;
    int chunk = (int)(p->ws_bytes / need);
    if (chunk < 1) { set_error("spike fast path: workspace too small"); return MVTB_EUNSUPPORTED; }
    if (chunk > n_volumes) chunk = n_volumes;
    // ~16 BraTS volumes' worth of voxels per round: enough CTAs to hide the two one-CTA-per-volume helper kernels,
    // and thousands of 2-D slices per round rather than 16 (a stack of 8192 240x240 slices is 4 rounds, not 512)
    {
        long long cap = (16LL * 240 * 240 * 155) / (long long)p->vol_real;
        if (cap < 16) cap = 16;
        if (cap > 65535) cap = 65535;                  // grid.y
        if (chunk > cap) chunk = (int)cap;
    }
    cf* w_partial = p->ws;
    cf* w_delta = w_partial + (size_t)chunk * MVTB_MAX_SPIKES * ctas;
    cf* w_tab = w_delta + (size_t)chunk * MVTB_MAX_SPIKES;
    const size_t smem = sizeof(cf) * (size_t)g.nmax * ((size_t)g.tlen + kSpRowsPerCta);
    for (int v0 = 0; v0 < n_volumes; v0 += chunk) {
        const int nv = n_volumes - v0 < chunk ? n_volumes - v0 : chunk;
        const SpVol* dvc = shared_desc ? dv : dv + v0;
        {
            ProfScope prof(p, MVTB_K_SPIKE_REDUCE, stream);
            MVTB_LAUNCH(k_spike_tables, dim3((unsigned)nv), dim3(kSpThreads), 0, stream, w_tab, g, dvc, shared_desc);
            MVTB_LAUNCH(k_spike_reduce, dim3((unsigned)ctas, (unsigned)nv), dim3(kSpThreads), smem, stream,
                        in + (size_t)v0 * p->vol_real, w_partial, g, dvc, shared_desc, (const cf*)w_tab);
        }
        {
            ProfScope prof(p, MVTB_K_SPIKE_APPLY, stream);
            MVTB_LAUNCH(k_spike_finalize, dim3((unsigned)nv), dim3(kSpThreads), 0, stream, (const cf*)w_partial, ctas, w_delta, g, dvc, shared_desc);
            MVTB_LAUNCH(k_spike_apply, dim3((unsigned)ctas, (unsigned)nv), dim3(kSpThreads), smem, stream,
                        in + (size_t)v0 * p->vol_real, out + (size_t)v0 * p->vol_real, (const cf*)w_delta, g, dvc, shared_desc,
                        (const cf*)w_tab, minmax_out, minmax_out ? vols_per_sample : 1, v0);
        }
    }
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}

int configure_spike_kernels(const mvtb_plan* p) {
#ifndef MVTB_EMU
    cudaDeviceProp prop;
    MVTB_CUDA(cudaGetDeviceProperties(&prop, p->device));
    cudaFuncAttributes a;
    MVTB_CUDA(cudaFuncGetAttributes(&a, k_spike_reduce));
    MVTB_CUDA(cudaFuncSetAttribute(k_spike_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - (int)a.sharedSizeBytes));
    MVTB_CUDA(cudaFuncGetAttributes(&a, k_spike_apply));
    MVTB_CUDA(cudaFuncSetAttribute(k_spike_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - (int)a.sharedSizeBytes));
#endif
    (void)p;
    return MVTB_OK;
}

}  // namespace mvtb
