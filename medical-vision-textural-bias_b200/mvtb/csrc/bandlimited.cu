// bandlimited.cu — the k-space chain when the mask keeps only a small ball of frequencies.
//
// RandFourierDiskMaskd (F:236-252) with the radii the scripts use (r = 9 ... 15; r = 12.5 in the
// 125/126/127 chains) keeps |f_d| <= F = floor(sqrt(thr)) on every axis, i.e. (2F+1)^2 (F+1) of the
// N_h N_w N_d/2 half-spectrum bins (8 125 of 4.5 M for 240x240x155, r = 12.5).  A full FFT computes
// 550x more bins than survive the mask.  This path computes only the surviving ones, as pruned
// DFTs with the symmetric-pair folding  x[h] +- x[N-h]  (cos part / sin part), in five kernels:
//
//   k_bl_fwd_h   x[v][H][W*D] real       -> Y[v][NF][W*D]      one thread per (w,d) column, streams the
//                                                             volume ONCE from HBM with coalesced loads
//   k_bl_fwd_w   Y[v][NF][W][D]          -> G[v][NF][K][D]     K = 2F+1
//   k_bl_mid     G: D-axis DFT to K bins, pointwise (mask / in-box spikes / wrap / 1/N), back
//   k_bl_inv_w   G[v][NF][K][D]          -> Y[v][NF][W][D]
//   k_bl_inv_h   Y                       -> out[v][H][W*D]     one thread per column, writes the volume
//                                                             ONCE; adds out-of-box spikes as plane waves
//                                                              and tracks per-sample min/max
//
// HBM traffic is the compulsory 8 B/voxel plus ~1 B/voxel of intermediates (Y is NF/H of the
// volume); arithmetic is ~2(2F+1) FMA per voxel pair and direction.  Results are the same
// numbers the general path produces (same pointwise stage, pointwise.cuh), to fp32 rounding.
#include <math.h>
#include <string.h>

#include "pointwise.cuh"

namespace mvtb {

int convert_desc(const mvtb_plan* p, const mvtb_chain_desc* u, DescDev* d);   // kspace_chain.cu

static const int kBlThreads = 256;

struct BlGeom {
    int H, W, D;
    int F;                    // kept |f| <= F on every axis; NF = F+1 rows of the h half-spectrum
    long long NC;             // W*D
    const float* tabC[3];     // [N][MVTB_BL_FT] cos(2 pi f n / N), axis 0 = D, 1 = W, 2 = H
    const float* tabS[3];
    const cf* twD;            // exp(-2 pi i t / D)
    float scale;              // 1/(H W D)
};

struct PlaneWave { int fh, fw, fd; float amp; };      // signed frequencies, amplitude incl. wrap weight and 1/N
struct PwPack {
    int n[MVTB_DESC_PACK];
    PlaneWave pw[MVTB_DESC_PACK][MVTB_BL_MAX_PW];
};

// cos / sin rows of one axis for n = 0 .. N/2 into shared memory, NFP floats per row
template <int NF>
__device__ __forceinline__ void bl_load_table(float* sc, float* ss, const float* __restrict__ tabC,
                                              const float* __restrict__ tabS, int N, int tid, int nthr) {
    constexpr int NFP = (NF + 3) & ~3;
    const int rows = N / 2 + 1;
    for (int e = tid; e < rows * NFP; e += nthr) {
        const int n = e / NFP, f = e - n * NFP;
        sc[e] = f < NF ? __ldg(tabC + n * MVTB_BL_FT + f) : 0.f;
        ss[e] = f < NF ? __ldg(tabS + n * MVTB_BL_FT + f) : 0.f;
    }
}

// ------------------------------------------------------------------ H axis forward: real -> NF complex rows
template <int NF>
__global__ void __launch_bounds__(256)
k_bl_fwd_h(const float* __restrict__ x, cf* __restrict__ Y, BlGeom g, int n_cblocks) {
    constexpr int NFP = (NF + 3) & ~3;
    MVTB_DYN_SMEM(smem_raw);
    float* sc = (float*)smem_raw;
    float* ss = sc + (g.H / 2 + 1) * NFP;
    const int tid = threadIdx.x;
    bl_load_table<NF>(sc, ss, g.tabC[2], g.tabS[2], g.H, tid, blockDim.x);
    __syncthreads();

    const long long vol = blockIdx.x / n_cblocks;
    const long long c = (long long)(blockIdx.x - vol * n_cblocks) * blockDim.x + tid;
    if (c >= g.NC) return;
    const float* xv = x + vol * g.H * g.NC + c;
    float re[NF], im[NF];
    const float x0 = xv[0];
    MVTB_UNROLL
    for (int f = 0; f < NF; ++f) { re[f] = x0; im[f] = 0.f; }
    const int H = g.H;
    if ((H & 1) == 0) {
        const float xn = xv[(long long)(H / 2) * g.NC];
        MVTB_UNROLL
        for (int f = 0; f < NF; ++f) re[f] += (f & 1) ? -xn : xn;
    }
    const int npair = (H - 1) / 2;
    int h = 1;
    for (; h + 3 <= npair; h += 4) {          // 8 independent loads in flight per thread
        float a[4], b[4];
        MVTB_UNROLL
        for (int u = 0; u < 4; ++u) {
            a[u] = xv[(long long)(h + u) * g.NC];
            b[u] = xv[(long long)(H - h - u) * g.NC];
        }
        MVTB_UNROLL
        for (int u = 0; u < 4; ++u) {
            const float e = a[u] + b[u], o = a[u] - b[u];
            const float* c_ = sc + (h + u) * NFP;
            const float* s_ = ss + (h + u) * NFP;
            MVTB_UNROLL
            for (int f = 0; f < NF; ++f) { re[f] = fmaf(e, c_[f], re[f]); im[f] = fmaf(-o, s_[f], im[f]); }
        }
    }
    for (; h <= npair; ++h) {
        const float a = xv[(long long)h * g.NC], b = xv[(long long)(H - h) * g.NC];
        const float e = a + b, o = a - b;
        const float* c_ = sc + h * NFP;
        const float* s_ = ss + h * NFP;
        MVTB_UNROLL
        for (int f = 0; f < NF; ++f) { re[f] = fmaf(e, c_[f], re[f]); im[f] = fmaf(-o, s_[f], im[f]); }
    }
    cf* yv = Y + vol * NF * g.NC + c;
    MVTB_UNROLL
    for (int f = 0; f < NF; ++f) yv[(long long)f * g.NC] = cmk(re[f], im[f]);
}

// ------------------------------------------------------------------ W axis forward: Y[NF][W][D] -> G[NF][K][D]
template <int NF>
__global__ void __launch_bounds__(256)
k_bl_fwd_w(const cf* __restrict__ Y, cf* __restrict__ G, BlGeom g, int n_tblocks) {
    constexpr int NFP = (NF + 3) & ~3;
    MVTB_DYN_SMEM(smem_raw);
    float* sc = (float*)smem_raw;
    float* ss = sc + (g.W / 2 + 1) * NFP;
    const int tid = threadIdx.x;
    bl_load_table<NF>(sc, ss, g.tabC[1], g.tabS[1], g.W, tid, blockDim.x);
    __syncthreads();

    const int W = g.W, D = g.D, K = 2 * g.F + 1;
    const long long vol = blockIdx.x / n_tblocks;
    const int t = (int)(blockIdx.x - vol * n_tblocks) * blockDim.x + tid;     // (fh, d)
    if (t >= NF * D) return;
    const int fh = t / D, d = t - fh * D;
    const cf* yv = Y + ((vol * NF + fh) * (long long)W) * D + d;
    cf P[NF], Q[NF];
    const cf y0 = yv[0];
    MVTB_UNROLL
    for (int f = 0; f < NF; ++f) { P[f] = y0; Q[f] = cmk(0.f, 0.f); }
    if ((W & 1) == 0) {
        const cf yn = yv[(long long)(W / 2) * D];
        MVTB_UNROLL
        for (int f = 0; f < NF; ++f) { P[f].x += (f & 1) ? -yn.x : yn.x; P[f].y += (f & 1) ? -yn.y : yn.y; }
    }
    const int npair = (W - 1) / 2;
    for (int w = 1; w <= npair; ++w) {
        const cf a = yv[(long long)w * D], b = yv[(long long)(W - w) * D];
        const cf e = cadd(a, b), o = csub(a, b);
        const float* c_ = sc + w * NFP;
        const float* s_ = ss + w * NFP;
        MVTB_UNROLL
        for (int f = 0; f < NF; ++f) {
            P[f].x = fmaf(e.x, c_[f], P[f].x); P[f].y = fmaf(e.y, c_[f], P[f].y);
            Q[f].x = fmaf(o.x, s_[f], Q[f].x); Q[f].y = fmaf(o.y, s_[f], Q[f].y);
        }
    }
    // X(+f) = P - iQ, X(-f) = P + iQ;  row j of G holds fw = j - F
    cf* gv = G + ((vol * NF + fh) * (long long)K) * D + d;
    MVTB_UNROLL
    for (int f = 0; f < NF; ++f) {
        if (f > g.F) break;
        gv[(long long)(g.F + f) * D] = cmk(P[f].x + Q[f].y, P[f].y - Q[f].x);
        if (f > 0) gv[(long long)(g.F - f) * D] = cmk(P[f].x - Q[f].y, P[f].y + Q[f].x);
    }
}

// ------------------------------------------------------------------ D axis both ways + pointwise; CTA = (vol, fh)
__global__ void __launch_bounds__(256)
k_bl_mid(cf* __restrict__ G, BlGeom g, int NF, DescPack pack) {
    MVTB_DYN_SMEM(smem_raw);
    const int D = g.D, K = 2 * g.F + 1, F = g.F;
    cf* sg = (cf*)smem_raw;            // [K][D]
    cf* sb = sg + K * D;               // [K][K]
    cf* st = sb + K * K;               // [D] exp(-2 pi i t / D)
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int vol = blockIdx.x / NF, fh = blockIdx.x - vol * NF;
    cf* gv = G + ((long long)vol * NF + fh) * K * D;
    for (int e = tid; e < K * D; e += nthr) sg[e] = gv[e];
    for (int e = tid; e < D; e += nthr) st[e] = __ldg(g.twD + e);
    __syncthreads();

    const DescDev& dsc = pack.d[pack.n == 1 ? 0 : vol];
    int shape[3];
    shape[0] = g.D; shape[1] = g.W; shape[2] = g.H;
    // B[jw][jd] = sum_d G[jw][d] exp(-2 pi i fd d / D), then the pointwise stage on that bin
    for (int o = tid; o < K * K; o += nthr) {
        const int jw = o / K, jd = o - jw * K;
        const int fd = jd - F;
        const int step = ((fd % D) + D) % D;
        int idx = 0;
        cf acc = cmk(0.f, 0.f);
        const cf* row = sg + jw * D;
        for (int d = 0; d < D; ++d) {
            const cf a = row[d], w = st[idx];
            acc.x = fmaf(a.x, w.x, fmaf(-a.y, w.y, acc.x));
            acc.y = fmaf(a.x, w.y, fmaf(a.y, w.x, acc.y));
            idx += step;
            if (idx >= D) idx -= D;
        }
        int ish[3];
        ish[0] = fd + D / 2;
        ish[1] = (jw - F) + g.W / 2;
        ish[2] = fh + g.H / 2;
        sb[o] = pointwise_bin(dsc, 3, shape, ish, acc, g.scale);
    }
    __syncthreads();
    // G'[jw][d] = sum_jd B[jw][jd] exp(+2 pi i fd d / D)
    for (int o = tid; o < K * D; o += nthr) {
        const int jw = o / D, d = o - jw * D;
        const cf* brow = sb + jw * K;
        // fd runs -F..F: start at (-F d) mod D and advance by d
        int idx = (int)((((long long)(-F) * d) % D + D) % D);
        cf acc = cmk(0.f, 0.f);
        for (int jd = 0; jd < K; ++jd) {
            const cf b = brow[jd], w = st[idx];          // conj(w) = exp(+...)
            acc.x = fmaf(b.x, w.x, fmaf(b.y, w.y, acc.x));
            acc.y = fmaf(b.y, w.x, fmaf(-b.x, w.y, acc.y));
            idx += d;
            if (idx >= D) idx -= D;
        }
        gv[o] = acc;
    }
}

// ------------------------------------------------------------------ W axis inverse: G[NF][K][D] -> Y[NF][W][D]
template <int NF>
__global__ void __launch_bounds__(256)
k_bl_inv_w(const cf* __restrict__ G, cf* __restrict__ Y, BlGeom g, int n_tblocks) {
    constexpr int NFP = (NF + 3) & ~3;
    MVTB_DYN_SMEM(smem_raw);
    float* sc = (float*)smem_raw;
    float* ss = sc + (g.W / 2 + 1) * NFP;
    const int tid = threadIdx.x;
    bl_load_table<NF>(sc, ss, g.tabC[1], g.tabS[1], g.W, tid, blockDim.x);
    __syncthreads();

    const int W = g.W, D = g.D, K = 2 * g.F + 1;
    const long long vol = blockIdx.x / n_tblocks;
    const int t = (int)(blockIdx.x - vol * n_tblocks) * blockDim.x + tid;
    if (t >= NF * D) return;
    const int fh = t / D, d = t - fh * D;
    const cf* gv = G + ((vol * NF + fh) * (long long)K) * D + d;
    // S_f = G(+f) + G(-f), T_f = G(+f) - G(-f);  y[w] = P + iQ, y[W-w] = P - iQ with
    // P = sum S_f cos, Q = sum T_f sin
    cf S[NF], T[NF];
    MVTB_UNROLL
    for (int f = 0; f < NF; ++f) {
        if (f <= g.F) {
            const cf gp = gv[(long long)(g.F + f) * D];
            const cf gm = f > 0 ? gv[(long long)(g.F - f) * D] : cmk(0.f, 0.f);
            S[f] = cadd(gp, gm);
            T[f] = f > 0 ? csub(gp, gm) : cmk(0.f, 0.f);
        } else {
            S[f] = cmk(0.f, 0.f);
            T[f] = cmk(0.f, 0.f);
        }
    }
    cf* yv = Y + ((vol * NF + fh) * (long long)W) * D + d;
    {
        cf s0 = cmk(0.f, 0.f), sn = cmk(0.f, 0.f);
        MVTB_UNROLL
        for (int f = 0; f < NF; ++f) {
            s0 = cadd(s0, S[f]);
            sn = (f & 1) ? csub(sn, S[f]) : cadd(sn, S[f]);
        }
        yv[0] = s0;
        if ((W & 1) == 0) yv[(long long)(W / 2) * D] = sn;
    }
    const int npair = (W - 1) / 2;
    for (int w = 1; w <= npair; ++w) {
        const float* c_ = sc + w * NFP;
        const float* s_ = ss + w * NFP;
        cf P = cmk(0.f, 0.f), Q = cmk(0.f, 0.f);
        MVTB_UNROLL
        for (int f = 0; f < NF; ++f) {
            P.x = fmaf(S[f].x, c_[f], P.x); P.y = fmaf(S[f].y, c_[f], P.y);
            Q.x = fmaf(T[f].x, s_[f], Q.x); Q.y = fmaf(T[f].y, s_[f], Q.y);
        }
        yv[(long long)w * D] = cmk(P.x - Q.y, P.y + Q.x);
        yv[(long long)(W - w) * D] = cmk(P.x + Q.y, P.y - Q.x);
    }
}

// ------------------------------------------------------------------ H axis inverse: Y[NF][W*D] -> real volume
__device__ __forceinline__ void bl_atomic_min(float* addr, float v) {
    v += 0.0f;
    if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v));
    else atomicMax((unsigned*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void bl_atomic_max(float* addr, float v) {
    v += 0.0f;
    if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
    else atomicMin((unsigned*)addr, __float_as_uint(v));
}

__device__ __forceinline__ void bl_unit(int f, int n, int N, float* c, float* s) {
    // exp(+2 pi i f n / N) with the integer product reduced mod N first
    long long m = ((long long)f * n) % N;
    if (m < 0) m += N;
    sincospif(2.0f * (float)m / (float)N, s, c);
}

template <int NF>
__global__ void __launch_bounds__(256)
k_bl_inv_h(const cf* __restrict__ Y, float* __restrict__ out, BlGeom g, int n_cblocks, PwPack pws, int pack_shared,
           float* __restrict__ minmax, int vols_per_sample, int vol_base) {
    constexpr int NFP = (NF + 3) & ~3;
    MVTB_DYN_SMEM(smem_raw);
    const int H = g.H;
    float* sc = (float*)smem_raw;
    float* ss = sc + (H / 2 + 1) * NFP;
    cf* seh = (cf*)(ss + (H / 2 + 1) * NFP);            // [MVTB_BL_MAX_PW][H/2+1] exp(+2 pi i fh h / H)
    const int tid = threadIdx.x;
    const int vol = blockIdx.x / n_cblocks;
    const int pslot = pack_shared ? 0 : vol;
    const int npw = pws.n[pslot];
    bl_load_table<NF>(sc, ss, g.tabC[2], g.tabS[2], H, tid, blockDim.x);
    for (int e = tid; e < npw * (H / 2 + 1); e += blockDim.x) {
        const int s = e / (H / 2 + 1), h = e - s * (H / 2 + 1);
        float c_, s_;
        bl_unit(pws.pw[pslot][s].fh, h, H, &c_, &s_);
        seh[e] = cmk(c_, s_);
    }
    __syncthreads();

    const long long c = (long long)(blockIdx.x - (long long)vol * n_cblocks) * blockDim.x + tid;
    float lo = __int_as_float(0x7f800000), hi = __int_as_float((int)0xff800000u);
    if (c < g.NC) {
        const cf* yv = Y + (long long)vol * NF * g.NC + c;
        float a[NF], b[NF];
        MVTB_UNROLL
        for (int f = 0; f < NF; ++f) {
            const cf y = yv[(long long)f * g.NC];
            const float cfw = f == 0 ? 1.f : 2.f;
            a[f] = cfw * y.x;
            b[f] = cfw * y.y;
        }
        // per-column factor of every plane wave: amp * exp(+2 pi i (fw w / W + fd d / D))
        cf E[MVTB_BL_MAX_PW];
        {
            const int w = (int)(c / g.D), d = (int)(c - (long long)w * g.D);
            MVTB_UNROLL
            for (int s = 0; s < MVTB_BL_MAX_PW; ++s) {
                if (s < npw) {
                    float cw, sw, cd, sd;
                    bl_unit(pws.pw[pslot][s].fw, w, g.W, &cw, &sw);
                    bl_unit(pws.pw[pslot][s].fd, d, g.D, &cd, &sd);
                    const float amp = pws.pw[pslot][s].amp;
                    E[s] = cmk(amp * (cw * cd - sw * sd), amp * (sw * cd + cw * sd));
                } else {
                    E[s] = cmk(0.f, 0.f);
                }
            }
        }
        float* ov = out + (long long)vol * H * g.NC + c;
        {
            float v0 = 0.f, vn = 0.f;
            MVTB_UNROLL
            for (int f = 0; f < NF; ++f) { v0 += a[f]; vn += (f & 1) ? -a[f] : a[f]; }
            MVTB_UNROLL
            for (int s = 0; s < MVTB_BL_MAX_PW; ++s) {
                if (s < npw) {
                    v0 += E[s].x;                               // exp(0) = 1
                    if ((H & 1) == 0) {
                        const cf eh = seh[s * (H / 2 + 1) + H / 2];
                        vn += E[s].x * eh.x - E[s].y * eh.y;
                    }
                }
            }
            ov[0] = v0;
            lo = fminf(lo, v0); hi = fmaxf(hi, v0);
            if ((H & 1) == 0) { ov[(long long)(H / 2) * g.NC] = vn; lo = fminf(lo, vn); hi = fmaxf(hi, vn); }
        }
        const int npair = (H - 1) / 2;
        for (int h = 1; h <= npair; ++h) {
            const float* c_ = sc + h * NFP;
            const float* s_ = ss + h * NFP;
            float P = 0.f, Q = 0.f;
            MVTB_UNROLL
            for (int f = 0; f < NF; ++f) { P = fmaf(a[f], c_[f], P); Q = fmaf(b[f], s_[f], Q); }
            MVTB_UNROLL
            for (int s = 0; s < MVTB_BL_MAX_PW; ++s) {
                if (s < npw) {
                    const cf eh = seh[s * (H / 2 + 1) + h];
                    P = fmaf(E[s].x, eh.x, P);
                    Q = fmaf(E[s].y, eh.y, Q);
                }
            }
            const float v1 = P - Q, v2 = P + Q;
            ov[(long long)h * g.NC] = v1;
            ov[(long long)(H - h) * g.NC] = v2;
            lo = fminf(lo, fminf(v1, v2));
            hi = fmaxf(hi, fmaxf(v1, v2));
        }
    }
    if (minmax != nullptr) {
        __shared__ float s_lo[32], s_hi[32];
        MVTB_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        const int lane = tid & 31, wp = tid >> 5, nw = (blockDim.x + 31) >> 5;
        if (lane == 0) { s_lo[wp] = lo; s_hi[wp] = hi; }
        __syncthreads();
        if (wp == 0) {
            lo = lane < nw ? s_lo[lane] : __int_as_float(0x7f800000);
            hi = lane < nw ? s_hi[lane] : __int_as_float((int)0xff800000u);
            MVTB_UNROLL
            for (int o = 16; o > 0; o >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0) {
                float* mm = minmax + 2 * ((vol_base + vol) / vols_per_sample);
                bl_atomic_min(mm, lo);
                bl_atomic_max(mm + 1, hi);
            }
        }
    }
}

// ------------------------------------------------------------------ host side
size_t bl_workspace_per_volume(const mvtb_plan* p, int F);

static int isqrt_ll(long long v) {
    long long r = (long long)floor(sqrt((double)v));
    while (r * r > v) --r;
    while ((r + 1) * (r + 1) <= v) ++r;
    return (int)r;
}

static int pick_nf(int need) {
    static const int avail[] = {4, 8, 13, 16};
    for (int a : avail)
        if (need <= a) return a;
    return 0;
}

// can this call take the band-limited path?  fills *F_out
bool bl_eligible(const mvtb_plan* p, const mvtb_chain_desc* desc, int n_desc, int* F_out) {
    if (p->ndim != 3 || p->opt_path == 1 || !p->bl_tab) return false;
    long long thr = desc[0].mask_thresh;
    for (int i = 0; i < n_desc; ++i) {
        const mvtb_chain_desc& d = desc[i];
        if (d.mask_kind != MVTB_MASK_DISK || d.mask_ndim != 3 || d.inside_off || d.mask_thresh < 0) return false;
        if (d.mask_thresh != thr) return false;
        if (d.wrap_naxes != 0 && d.wrap_naxes != 3) return false;
    }
    const int F = isqrt_ll(thr);
    const int nf = pick_nf(F + 1);
    if (nf == 0) return false;
    if (2 * nf - 1 > p->shape[2] || 2 * F + 1 > p->shape[1] || 2 * F + 1 > p->shape[0]) return false;
    for (int i = 0; i < n_desc; ++i) {                      // at most MVTB_BL_MAX_PW spikes outside the kept box
        int outside = 0;
        for (int s = 0; s < desc[i].n_spikes; ++s) {
            bool in_box = true;
            for (int a = 0; a < 3; ++a) {
                const int n = p->shape[2 - a];                // user order: outermost (H) first
                const int idx = desc[i].spikes[s].idx[a];
                if (idx < 0 || idx >= n) return false;      // let the general path report the error
                if (abs(idx - n / 2) > F) in_box = false;
            }
            if (!in_box) ++outside;
        }
        if (outside > MVTB_BL_MAX_PW) return false;
    }
    if (bl_workspace_per_volume(p, F) > p->ws_bytes) return false;
    *F_out = F;
    return true;
}

size_t bl_workspace_per_volume(const mvtb_plan* p, int F) {
    const int nf = pick_nf(F + 1);
    const size_t W = p->shape[1], D = p->shape[0], K = 2 * (size_t)F + 1;
    return sizeof(cf) * (size_t)nf * (W * D + K * D);
}

template <int NF>
static int bl_run(mvtb_plan* p, const float* in, float* out, int n_volumes, const mvtb_chain_desc* desc, int n_desc,
                  int F, float* minmax_out, int vols_per_sample, void* stream) {
    BlGeom g;
    g.D = p->shape[0]; g.W = p->shape[1]; g.H = p->shape[2];
    g.F = F;
    g.NC = (long long)g.W * g.D;
    for (int a = 0; a < 3; ++a) {
        g.tabC[a] = p->bl_tab + (size_t)p->bl_off[a];
        g.tabS[a] = g.tabC[a] + (size_t)p->shape[a] * MVTB_BL_FT;
    }
    g.twD = p->ax[0].tw;
    g.scale = (float)(1.0 / ((double)g.H * g.W * g.D));
    constexpr int NFP = (NF + 3) & ~3;
    const int K = 2 * F + 1;
    const size_t per_vol = bl_workspace_per_volume(p, F);
    int chunk = (int)(p->ws_bytes / per_vol);
    if (chunk < 1) { set_error("band-limited path: workspace too small"); return MVTB_EUNSUPPORTED; }
    if (chunk > n_volumes) chunk = n_volumes;
    const int n_cblocks = (int)((g.NC + kBlThreads - 1) / kBlThreads);
    const int n_tblocks = (NF * g.D + kBlThreads - 1) / kBlThreads;
    const size_t smem_h = sizeof(float) * 2 * (g.H / 2 + 1) * NFP;
    const size_t smem_w = sizeof(float) * 2 * (g.W / 2 + 1) * NFP;
    const size_t smem_hi = smem_h + sizeof(cf) * MVTB_BL_MAX_PW * (g.H / 2 + 1);
    const size_t smem_mid = sizeof(cf) * ((size_t)K * g.D + (size_t)K * K + g.D);

    for (int v0 = 0; v0 < n_volumes; v0 += chunk) {
        const int nv = n_volumes - v0 < chunk ? n_volumes - v0 : chunk;
        cf* Y = p->ws;
        cf* G = Y + (size_t)chunk * NF * g.NC;
        {
            ProfScope prof(p, MVTB_K_BL_FWD_H, stream);
            auto kern = k_bl_fwd_h<NF>;
            MVTB_LAUNCH(kern, dim3((unsigned)(n_cblocks * nv)), dim3(kBlThreads), smem_h, stream,
                        in + (size_t)v0 * p->vol_real, Y, g, n_cblocks);
        }
        {
            ProfScope prof(p, MVTB_K_BL_FWD_W, stream);
            auto kern = k_bl_fwd_w<NF>;
            MVTB_LAUNCH(kern, dim3((unsigned)(n_tblocks * nv)), dim3(kBlThreads), smem_w, stream, (const cf*)Y, G, g, n_tblocks);
        }
        // per-volume descriptors travel by value, MVTB_DESC_PACK volumes per launch
        const int sub = n_desc == 1 ? nv : MVTB_DESC_PACK;
        for (int w0 = 0; w0 < nv; w0 += sub) {
            const int nw = nv - w0 < sub ? nv - w0 : sub;
            DescPack pack;
            PwPack pws;
            memset(&pack, 0, sizeof(pack));
            memset(&pws, 0, sizeof(pws));
            pack.n = n_desc == 1 ? 1 : (nw == 1 ? 1 : nw);
            const int nslots = n_desc == 1 ? 1 : nw;
            for (int i = 0; i < nslots; ++i) {
                mvtb_chain_desc u = desc[n_desc == 1 ? 0 : v0 + w0 + i];
                // out-of-box spikes become plane waves in k_bl_inv_h; in-box ones stay in the pointwise stage
                mvtb_chain_desc inbox = u;
                inbox.n_spikes = 0;
                for (int s = 0; s < u.n_spikes; ++s) {
                    const int fh = u.spikes[s].idx[0] - g.H / 2, fw = u.spikes[s].idx[1] - g.W / 2, fd = u.spikes[s].idx[2] - g.D / 2;
                    const bool in_box = abs(fh) <= F && abs(fw) <= F && abs(fd) <= F;
                    if (in_box) {
                        inbox.spikes[inbox.n_spikes++] = u.spikes[s];
                    } else {
                        // the bin is masked to exactly 0, so new = amp * exp(i angle(0)) = amp (F:384-390)
                        float a = u.spikes[s].amplitude * g.scale;
                        if (u.wrap_naxes == 3)
                            for (int ax = 0; ax < 3; ++ax)
                                if (u.spikes[s].idx[ax] & 1) a *= u.wrap_alpha;
                        PlaneWave& pw = pws.pw[i][pws.n[i]++];
                        pw.fh = fh; pw.fw = fw; pw.fd = fd; pw.amp = a;
                    }
                }
                int rc = convert_desc(p, &inbox, &pack.d[i]);
                if (rc != MVTB_OK) return rc;
            }
            {
                ProfScope prof(p, MVTB_K_BL_MID, stream);
                MVTB_LAUNCH(k_bl_mid, dim3((unsigned)(NF * nw)), dim3(kBlThreads), smem_mid, stream,
                            G + (size_t)w0 * NF * K * g.D, g, NF, pack);
            }
            {
                ProfScope prof(p, MVTB_K_BL_INV_W, stream);
                auto kern = k_bl_inv_w<NF>;
                MVTB_LAUNCH(kern, dim3((unsigned)(n_tblocks * nw)), dim3(kBlThreads), smem_w, stream,
                            (const cf*)(G + (size_t)w0 * NF * K * g.D), Y + (size_t)w0 * NF * g.NC, g, n_tblocks);
            }
            {
                ProfScope prof(p, MVTB_K_BL_INV_H, stream);
                auto kern = k_bl_inv_h<NF>;
                MVTB_LAUNCH(kern, dim3((unsigned)(n_cblocks * nw)), dim3(kBlThreads), smem_hi, stream,
                            (const cf*)(Y + (size_t)w0 * NF * g.NC), out + (size_t)(v0 + w0) * p->vol_real, g, n_cblocks, pws,
                            n_desc == 1 ? 1 : 0, minmax_out, minmax_out ? vols_per_sample : 1, v0 + w0);
            }
        }
    }
    MVTB_CUDA(cudaGetLastError());
    return MVTB_OK;
}

int bl_chain(mvtb_plan* p, const float* in, float* out, int n_volumes, const mvtb_chain_desc* desc, int n_desc,
             int F, float* minmax_out, int vols_per_sample, void* stream) {
    switch (pick_nf(F + 1)) {
        case 4: return bl_run<4>(p, in, out, n_volumes, desc, n_desc, F, minmax_out, vols_per_sample, stream);
        case 8: return bl_run<8>(p, in, out, n_volumes, desc, n_desc, F, minmax_out, vols_per_sample, stream);
        case 13: return bl_run<13>(p, in, out, n_volumes, desc, n_desc, F, minmax_out, vols_per_sample, stream);
        case 16: return bl_run<16>(p, in, out, n_volumes, desc, n_desc, F, minmax_out, vols_per_sample, stream);
        default: set_error("band-limited path: F=%d not instantiated", F); return MVTB_EUNSUPPORTED;
    }
}

#ifndef MVTB_EMU
template <typename K>
static int bl_big_smem(K kern, int optin) {
    cudaFuncAttributes a;
    MVTB_CUDA(cudaFuncGetAttributes(&a, kern));
    MVTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)a.sharedSizeBytes));
    return MVTB_OK;
}
template <int NF>
static int bl_configure_nf(int optin) {
    int rc;
    if ((rc = bl_big_smem(k_bl_fwd_h<NF>, optin)) != MVTB_OK) return rc;
    if ((rc = bl_big_smem(k_bl_fwd_w<NF>, optin)) != MVTB_OK) return rc;
    if ((rc = bl_big_smem(k_bl_inv_w<NF>, optin)) != MVTB_OK) return rc;
    if ((rc = bl_big_smem(k_bl_inv_h<NF>, optin)) != MVTB_OK) return rc;
    return MVTB_OK;
}
#endif

int configure_bl_kernels(const mvtb_plan* p) {
#ifndef MVTB_EMU
    cudaDeviceProp prop;
    MVTB_CUDA(cudaGetDeviceProperties(&prop, p->device));
    const int optin = (int)prop.sharedMemPerBlockOptin;
    int rc;
    if ((rc = bl_configure_nf<4>(optin)) != MVTB_OK) return rc;
    if ((rc = bl_configure_nf<8>(optin)) != MVTB_OK) return rc;
    if ((rc = bl_configure_nf<13>(optin)) != MVTB_OK) return rc;
    if ((rc = bl_configure_nf<16>(optin)) != MVTB_OK) return rc;
    if ((rc = bl_big_smem(k_bl_mid, optin)) != MVTB_OK) return rc;
#endif
    (void)p;
    return MVTB_OK;
}

}  // namespace mvtb
