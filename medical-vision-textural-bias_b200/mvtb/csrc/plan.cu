// plan.cu — FFT plan: factorisation, twiddle / digit-reversal tables, workspace; error text.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "mvtb_common.cuh"

namespace mvtb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

int configure_chain_kernels(const mvtb_plan* p);   // kspace_chain.cu: opt in to large dynamic shared memory
int configure_bl_kernels(const mvtb_plan* p);      // bandlimited.cu
int configure_spike_kernels(const mvtb_plan* p);   // spike_fast.cu

// prime factors <= 31, twos paired into fours, ascending (the radix-31 stage comes last,
// where the DIF stage has no twiddle multiplies)
static bool factorise(int n, std::vector<int>& out) {
    static const int primes[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31};
    int twos = 0;
    std::vector<int> odd;
    for (int p : primes) {
        while (n % p == 0) {
            if (p == 2) ++twos; else odd.push_back(p);
            n /= p;
        }
    }
    for (int q = 37; n > 1; q += 2) {                       // remaining prime factors > 31: generic direct-DFT stages
        if ((long long)q * q > n) { odd.push_back(n); n = 1; break; }
        while (n % q == 0) { odd.push_back(q); n /= q; }
    }
    out.clear();
    if (twos & 1) out.push_back(2);
    for (int i = 0; i < twos / 2; ++i) out.push_back(4);
    // keep ascending order overall
    std::vector<int> all(out);
    all.insert(all.end(), odd.begin(), odd.end());
    for (size_t i = 1; i < all.size(); ++i)
        for (size_t j = i; j > 0 && all[j - 1] > all[j]; --j) { int t = all[j]; all[j] = all[j - 1]; all[j - 1] = t; }
    out = all;
    return (int)out.size() <= MVTB_MAX_STAGES;
}

// Copies a small host array to device memory on `stream` without blocking the host: the bytes go
// through one of MVTB_STAGE_SLOTS pinned slots; a slot is reused only after the copy that last read
// it has executed.  The device copy is valid for kernels enqueued on the same stream after this call.
int plan_stage_upload(mvtb_plan* p, const void* src, size_t bytes, void* stream, void** dptr) {
    const int s = p->stage_next;
    p->stage_next = (s + 1) % MVTB_STAGE_SLOTS;
    if (p->stage_ev[s]) MVTB_CUDA(cudaEventSynchronize(p->stage_ev[s]));
    if (bytes > p->stage_cap[s]) {
        if (p->stage_h[s]) cudaFreeHost(p->stage_h[s]);
        if (p->stage_d[s]) cudaFree(p->stage_d[s]);       // cudaFree waits for work that may still read it
        p->stage_h[s] = nullptr; p->stage_d[s] = nullptr; p->stage_cap[s] = 0;
        size_t cap = bytes < 16384 ? 16384 : 2 * bytes;
        MVTB_CUDA(cudaMallocHost(&p->stage_h[s], cap));
        MVTB_CUDA(cudaMalloc(&p->stage_d[s], cap));
        p->stage_cap[s] = cap;
    }
    if (!p->stage_ev[s]) MVTB_CUDA(cudaEventCreate(&p->stage_ev[s]));
    memcpy(p->stage_h[s], src, bytes);
    MVTB_CUDA(cudaMemcpyAsync(p->stage_d[s], p->stage_h[s], bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    MVTB_CUDA(cudaEventRecord(p->stage_ev[s], (cudaStream_t)stream));
    *dptr = p->stage_d[s];
    return MVTB_OK;
}

}  // namespace mvtb

using namespace mvtb;

extern "C" int mvtb_version(void) { return MVTB_VERSION; }

extern "C" int mvtb_last_error(char* buf, int n) {
    int len = (int)strlen(g_err);
    if (buf && n > 0) {
        int c = len < n - 1 ? len : n - 1;
        memcpy(buf, g_err, c);
        buf[c] = 0;
    }
    return len;
}

extern "C" int mvtb_plan_create(mvtb_plan** out, int ndim_fft, const int* fft_shape, int chunk_volumes, int device) {
    if (!out || !fft_shape) { set_error("plan_create: null argument"); return MVTB_EINVAL; }
    *out = nullptr;
    if (ndim_fft < 2 || ndim_fft > MVTB_MAX_FFT_DIMS) { set_error("plan_create: ndim_fft=%d not in 2..4", ndim_fft); return MVTB_EINVAL; }
    if (chunk_volumes < 1) { set_error("plan_create: chunk_volumes must be >= 1"); return MVTB_EINVAL; }
    for (int i = 0; i < ndim_fft; ++i)
        if (fft_shape[i] < 1) { set_error("plan_create: axis %d has length %d", i, fft_shape[i]); return MVTB_EINVAL; }

#ifndef MVTB_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        set_error("plan_create: CUDA device %d not available (%d visible); libmvtb has no CPU path", device, ndev);
        return MVTB_ENODEVICE;
    }
#endif
    MVTB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MVTB_CUDA(cudaGetDeviceProperties(&prop, device));

    mvtb_plan* p = (mvtb_plan*)calloc(1, sizeof(mvtb_plan));
    if (!p) return MVTB_ENOMEM;
    // Leading FFT axes of length 1 (the channel axis of a (B,1,H,W,D) batch under GibbsNoiseLayer's "rank - 1"
    // rule, S:81) are length-1 DFTs: identity, zero mask term, even fftshift-ed index.  Drop them so that such
    // a 4-D transform runs as the 3-D one it is; the chain entry shifts the descriptors accordingly.
    int lead = 0;
    while (ndim_fft - lead > 2 && fft_shape[lead] == 1) ++lead;
    p->lead_drop = lead;
    fft_shape += lead;
    ndim_fft -= lead;
    p->ndim = ndim_fft;
    p->opt_quad = 1;
    p->opt_async = getenv("MVTB_NO_ASYNC") ? 0 : 1;
    p->opt_fusemid = getenv("MVTB_NO_FUSEMID") ? 0 : 1;
    p->opt_fusesp = getenv("MVTB_NO_FUSESP") ? 0 : 1;
    p->opt_tc = getenv("MVTB_TC") ? atoi(getenv("MVTB_TC")) : 1;
    p->opt_tc_inv = getenv("MVTB_TC_INV") ? atoi(getenv("MVTB_TC_INV")) : 0;
    p->opt_bits_overlap = getenv("MVTB_BITS_OVERLAP") ? atoi(getenv("MVTB_BITS_OVERLAP")) : 1;
    p->tci_par_vols = getenv("MVTB_TCI_PV") ? atoi(getenv("MVTB_TCI_PV")) : 4;
    if (p->tci_par_vols < 1) p->tci_par_vols = 1;
    p->tc_tma = getenv("MVTB_TC_TMA") ? atoi(getenv("MVTB_TC_TMA")) : 0;
    p->is_chunk = getenv("MVTB_IS_CHUNK") ? atoi(getenv("MVTB_IS_CHUNK")) : 0;
    p->is_hs = getenv("MVTB_IS_HS") ? atoi(getenv("MVTB_IS_HS")) : 2;
    p->is_lag = getenv("MVTB_IS_LAG") ? atoi(getenv("MVTB_IS_LAG")) : -1;
    p->is_spread_pct = getenv("MVTB_IS_SPREAD") ? atoi(getenv("MVTB_IS_SPREAD")) : 100;
    p->is_max_sample_mb = getenv("MVTB_IS_MAX_MB") ? atoi(getenv("MVTB_IS_MAX_MB")) : 48;
    p->is_store = getenv("MVTB_IS_STORE") ? atoi(getenv("MVTB_IS_STORE")) : 2;
    p->chunk = chunk_volumes;
    p->device = device;
    p->num_sms = prop.multiProcessorCount;
    for (int a = 0; a < ndim_fft; ++a) p->shape[a] = fft_shape[ndim_fft - 1 - a];   // axis 0 = last
    p->nh = p->shape[0] / 2 + 1;
    p->vol_real = 1;
    p->vol_half = (size_t)p->nh;
    for (int a = 0; a < ndim_fft; ++a) p->vol_real *= (size_t)p->shape[a];
    for (int a = 1; a < ndim_fft; ++a) p->vol_half *= (size_t)p->shape[a];

    // ---- tables (one host image, one device allocation)
    std::vector<std::vector<int>> radices(ndim_fft);
    size_t bytes = 0;
    for (int a = 0; a < ndim_fft; ++a) {
        if (!factorise(p->shape[a], radices[a])) {
            set_error("plan_create: axis length %d needs more than %d radix stages", p->shape[a], MVTB_MAX_STAGES);
            free(p);
            return MVTB_EUNSUPPORTED;
        }
        bytes += (size_t)p->shape[a] * (sizeof(cf) + 2 * sizeof(int));
        bytes = (bytes + 15) & ~(size_t)15;
    }
    std::vector<unsigned char> host(bytes);
    void* dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, bytes);
    if (e != cudaSuccess) { free(p); return cuda_fail(e, "cudaMalloc(tables)"); }
    size_t off = 0;
    for (int a = 0; a < ndim_fft; ++a) {
        const int n = p->shape[a];
        AxisDev& ax = p->ax[a];
        ax.n = n;
        ax.nstage = (int)radices[a].size();
        ax.generic = 0;
        for (int s = 0; s < ax.nstage; ++s) { ax.radix[s] = radices[a][s]; if (ax.radix[s] > 31) ax.generic = 1; }
        // greedy pairing of consecutive small radices (product <= 20) into register-fused passes
        for (int s = 0; s < MVTB_MAX_STAGES; ++s) ax.fuse[s] = 0;
        if (!getenv("MVTB_NO_FUSE"))
            for (int s = 0; s + 1 < ax.nstage; ++s) {
                const int r1 = ax.radix[s], r2 = ax.radix[s + 1];
                if (r1 <= 5 && r2 <= 5 && r1 * r2 <= 20 && r1 * r2 != 4) { ax.fuse[s] = 1; ++s; }
            }
        ax.npass = 0;
        for (int s = 0, L = n; s < ax.nstage; ++s) {
            PassDev& ps = ax.pass[ax.npass++];
            ps.r1 = ax.radix[s];
            ps.r2 = ax.fuse[s] ? ax.radix[s + 1] : 0;
            const int R = ps.r1 * (ps.r2 ? ps.r2 : 1);
            ps.L = L;
            ps.m = L / R;
            ps.per_seq = n / R;
            ps.ts1 = n / L;
            ps.ts2 = ps.r2 ? n / (L / ps.r1) : 0;
            ps.magic_m = ps.m > 1 ? 0xFFFFFFFFu / (unsigned)ps.m + 1u : 0u;
            ps.magic_ps = ps.per_seq > 1 ? 0xFFFFFFFFu / (unsigned)ps.per_seq + 1u : 0u;
            L /= R;
            if (ps.r2) ++s;
        }
        cf* tw = (cf*)(host.data() + off);
        ax.tw = (const cf*)((unsigned char*)dev + off);
        for (int t = 0; t < n; ++t) {
            double ang = -2.0 * M_PI * (double)t / (double)n;
            tw[t].x = (float)cos(ang);
            tw[t].y = (float)sin(ang);
        }
        off += (size_t)n * sizeof(cf);
        int* pos2k = (int*)(host.data() + off);
        ax.pos2k = (const int*)((unsigned char*)dev + off);
        off += (size_t)n * sizeof(int);
        int* k2pos = (int*)(host.data() + off);
        ax.k2pos = (const int*)((unsigned char*)dev + off);
        off += (size_t)n * sizeof(int);
        off = (off + 15) & ~(size_t)15;
        for (int pos = 0; pos < n; ++pos) {
            int rem = pos, L = n, k = 0, mult = 1;
            for (int s = 0; s < ax.nstage; ++s) {
                const int R = ax.radix[s], m = L / R, q = rem / m;
                rem -= q * m;
                k += q * mult;
                mult *= R;
                L = m;
            }
            pos2k[pos] = k;
            k2pos[k] = pos;
        }
    }
    e = cudaMemcpy(dev, host.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); free(p); return cuda_fail(e, "cudaMemcpy(tables)"); }
    p->table_mem = dev;

    // ---- kernel geometry
    // Row pitch of the rows kernels' tile (complex elements): at least 2 nh (the inverse kernel stages a pair of
    // half-spectrum rows in one tile row), and among the next 32 values the one with the fewest shared-memory bank
    // conflicts: for every pass, the 8-byte accesses of the first 128 tasks are binned per half-warp into the 16
    // bank pairs; a half-warp costs its fullest bin.  (155 = 5 x 31: pitch 157 put consecutive sequences 13 bank
    // pairs apart, so the 5-wide groups of the radix-31 pass overlapped two by two.)
    int pitch = 2 * p->nh;
    {
        const AxisDev& ax0 = p->ax[0];
        long long best = -1;
        for (int cand = 2 * p->nh; cand < 2 * p->nh + 32; ++cand) {
            long long cost = 0;
            for (int s = 0; s < ax0.npass; ++s) {
                const PassDev& ps = ax0.pass[s];
                const int R = ps.r1 * (ps.r2 ? ps.r2 : 1);
                long long pass_cost = 0;
                for (int half = 0; half < 8; ++half) {
                    int bins[16] = {0};
                    int worst = 0;
                    for (int l = 0; l < 16; ++l) {
                        const int task = half * 16 + l;
                        const int sq = task / ps.per_seq, b = task % ps.per_seq;
                        const int blk = b / ps.m, j = b % ps.m;
                        const int addr = sq * cand + blk * ps.L + j;
                        const int v = ++bins[addr & 15];
                        worst = v > worst ? v : worst;
                    }
                    pass_cost += worst;
                }
                cost += pass_cost * R;
            }
            if (best < 0 || cost < best) { best = cost; pitch = cand; }
        }
    }
    p->row_pitch = pitch;
    const size_t smem_cap = (size_t)prop.sharedMemPerBlockOptin;
    const size_t gen0 = p->ax[0].generic ? 2 : 1;              // generic stages need a scratch copy of the tile
    size_t per_pair = (size_t)pitch * sizeof(cf) * gen0;
    // Row pairs per CTA (kernels run 128 threads): at most ~36 KB of tile so that several CTAs share an SM, and
    // among those sizes the one that wastes the fewest thread slots in the stage passes, weighting a pass by its
    // radix (5 x 31 = 155: the radix-31 pass has 5 butterflies per pair, so 25 pairs fill 125 of 128 threads
    // where 26 would need a second, almost empty round).
    int rp_cap = (int)((size_t)36 * 1024 / per_pair);
    if (rp_cap > 64) rp_cap = 64;
    if (rp_cap < 1) rp_cap = 1;
    int rp = rp_cap;
    {
        const AxisDev& ax0 = p->ax[0];
        double best = 1e300;
        for (int c = rp_cap; c >= (rp_cap + 1) / 2; --c) {
            double cost = 0.0;
            for (int s = 0; s < ax0.nstage; ++s) {
                int R = ax0.radix[s];
                double per_task = R > 5 ? (double)R * R : 4.0 * R;      // direct prime butterfly: R^2; small radix: ~R log R
                if (ax0.fuse[s]) { R *= ax0.radix[s + 1]; ++s; per_task = 4.0 * R; }
                const long long tasks = (long long)(ax0.n / R) * c;
                cost += (double)((tasks + 127) / 128) * per_task;
            }
            cost /= (double)c;
            if (cost < best - 1e-12) { best = cost; rp = c; }
        }
    }
    if (per_pair > smem_cap) {
        set_error("plan_create: last axis of length %d does not fit in shared memory", p->shape[0]);
        cudaFree(dev); free(p);
        return MVTB_EUNSUPPORTED;
    }
    p->rows_pairs_per_cta = rp;
    p->axis_tile = getenv("MVTB_AXIS_TILE") ? atoi(getenv("MVTB_AXIS_TILE")) : 16;
    if (p->axis_tile != 4 && p->axis_tile != 8 && p->axis_tile != 16 && p->axis_tile != 32) p->axis_tile = 16;
    for (int a = 1; a < ndim_fft; ++a) {
        const size_t gen = p->ax[a].generic ? 2 : 1;
        const size_t tabb = (size_t)p->shape[a] * 16 + 16;      // the outermost axis also holds the pointwise stage's per-bin table
        while (p->axis_tile > 1 && (size_t)p->shape[a] * p->axis_tile * sizeof(cf) * gen + tabb > smem_cap) p->axis_tile /= 2;
        if ((size_t)p->shape[a] * p->axis_tile * sizeof(cf) * gen + tabb > smem_cap) {
            set_error("plan_create: axis of length %d does not fit in shared memory", p->shape[a]);
            cudaFree(dev); free(p);
            return MVTB_EUNSUPPORTED;
        }
    }

    // ---- workspace
    p->ws_bytes = (size_t)chunk_volumes * p->vol_half * sizeof(cf);
    e = cudaMalloc((void**)&p->ws, p->ws_bytes);
    if (e != cudaSuccess) { cudaFree(dev); free(p); return cuda_fail(e, "cudaMalloc(workspace)"); }

    // ---- band-limited path tables: cos/sin(2 pi f n / N) for f < MVTB_BL_FT, per axis (3-D plans)
    if (ndim_fft == 3) {
        size_t nfl = 0;
        for (int a = 0; a < 3; ++a) { p->bl_off[a] = nfl; nfl += (size_t)2 * p->shape[a] * MVTB_BL_FT; }
        std::vector<float> tab(nfl);
        for (int a = 0; a < 3; ++a) {
            const int n = p->shape[a];
            float* tc = tab.data() + p->bl_off[a];
            float* ts = tc + (size_t)n * MVTB_BL_FT;
            for (int i = 0; i < n; ++i)
                for (int f = 0; f < MVTB_BL_FT; ++f) {
                    const long long m = ((long long)f * i) % n;
                    const double ang = 2.0 * M_PI * (double)m / (double)n;
                    tc[(size_t)i * MVTB_BL_FT + f] = (float)cos(ang);
                    ts[(size_t)i * MVTB_BL_FT + f] = (float)sin(ang);
                }
        }
        e = cudaMalloc((void**)&p->bl_tab, nfl * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(p->bl_tab, tab.data(), nfl * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(p->ws); cudaFree(dev); free(p); return cuda_fail(e, "band-limited tables"); }
    }

    int rc = configure_chain_kernels(p);
    if (rc == MVTB_OK) rc = configure_bl_kernels(p);
    if (rc == MVTB_OK) rc = configure_spike_kernels(p);
    if (rc != MVTB_OK) { cudaFree(p->ws); cudaFree(dev); free(p); return rc; }
    *out = p;
    return MVTB_OK;
}

extern "C" int mvtb_plan_destroy(mvtb_plan* p) {
    if (!p) return MVTB_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    if (p->ws) cudaFree(p->ws);
    if (p->table_mem) cudaFree(p->table_mem);
    if (p->bl_tab) cudaFree(p->bl_tab);
    if (p->bl_ws) cudaFree(p->bl_ws);
    if (p->is_sync) cudaFree(p->is_sync);
    for (int i = 0; i < 8; ++i) {
        if (p->tc_tab_fwd[i]) cudaFree(p->tc_tab_fwd[i]);
        if (p->tc_tab_inv[i]) cudaFree(p->tc_tab_inv[i]);
    }
    if (p->tc_status) cudaFree(p->tc_status);
    if (p->tc_status_h) cudaFreeHost(p->tc_status_h);
    if (p->tc_bits) cudaFree(p->tc_bits);
#ifndef MVTB_EMU
    if (p->side_stream) { cudaStreamDestroy(p->side_stream); cudaEventDestroy(p->ev_fork); cudaEventDestroy(p->ev_join); }
#endif
    for (int s = 0; s < MVTB_STAGE_SLOTS; ++s) {
        if (p->stage_h[s]) cudaFreeHost(p->stage_h[s]);
        if (p->stage_d[s]) cudaFree(p->stage_d[s]);
        if (p->stage_ev[s]) cudaEventDestroy(p->stage_ev[s]);
    }
    if (p->prof_ev[0])
        for (int i = 0; i < 2 * MVTB_PROF_MAX; ++i) cudaEventDestroy(p->prof_ev[i]);
    free(p);
    return MVTB_OK;
}

extern "C" unsigned long long mvtb_launch_count(void) { return g_launches.load(); }

extern "C" const char* mvtb_kernel_name(int kind) {
    static const char* names[MVTB_K_KINDS] = {"k_rows_fwd", "k_axis<FWD>", "k_axis<MID>", "k_axis<INV>", "k_rows_inv",
                                              "k_bl_fwd_h", "k_bl_fwd_w", "k_bl_mid", "k_bl_inv_w", "k_bl_inv_h", "k_spike_reduce", "k_spike_apply", "k_rows_wrap", "k_bl_inv_sp", "k_bl_fwd_tc", "k_bl_inv_tc", "k_bl_mm_tc", "k_sp_bits"};
    return (kind >= 0 && kind < MVTB_K_KINDS) ? names[kind] : "";
}

extern "C" int mvtb_plan_set_path(mvtb_plan* p, int path) {
    if (!p || path < MVTB_PATH_AUTO || path > MVTB_PATH_BL_TC) { set_error("plan_set_path: bad argument"); return MVTB_EINVAL; }
    p->opt_tc = path == MVTB_PATH_BL_TC ? 1 : (path == MVTB_PATH_BL_CUDACORE ? 0 : (getenv("MVTB_TC") ? atoi(getenv("MVTB_TC")) : 1));
    p->opt_tc_inv = path == MVTB_PATH_BL_TC ? 1 : (getenv("MVTB_TC_INV") ? atoi(getenv("MVTB_TC_INV")) : 0);
    p->opt_path = path == MVTB_PATH_GENERAL ? MVTB_PATH_GENERAL : MVTB_PATH_AUTO;
    p->opt_quad = path == MVTB_PATH_BL_PAIRS ? 0 : 1;
    p->opt_fusemid = (path == MVTB_PATH_BL_SPLIT || getenv("MVTB_NO_FUSEMID")) ? 0 : 1;
    return MVTB_OK;
}

extern "C" int mvtb_plan_tc_status(mvtb_plan* p) {
    if (!p) { set_error("plan_tc_status: null plan"); return MVTB_EINVAL; }
    if (!p->tc_status) return 0;
    MVTB_CUDA(cudaSetDevice(p->device));
    int st = 0;
    MVTB_CUDA(cudaMemcpy(&st, p->tc_status, sizeof(int), cudaMemcpyDeviceToHost));
    return st;
}

extern "C" int mvtb_plan_profile(mvtb_plan* p, int enable) {
    if (!p) { set_error("plan_profile: null plan"); return MVTB_EINVAL; }
    MVTB_CUDA(cudaSetDevice(p->device));
    if (enable) {
        if (!p->prof_ev[0])
            for (int i = 0; i < 2 * MVTB_PROF_MAX; ++i) MVTB_CUDA(cudaEventCreate(&p->prof_ev[i]));
        p->prof_n = 0;
        memset(p->prof_ms, 0, sizeof(p->prof_ms));
        memset(p->prof_cnt, 0, sizeof(p->prof_cnt));
    }
    p->profiling = enable ? 1 : 0;
    return MVTB_OK;
}

extern "C" int mvtb_plan_profile_read(mvtb_plan* p, double* ms_sum, int* counts) {
    if (!p || !ms_sum || !counts) { set_error("plan_profile_read: null argument"); return MVTB_EINVAL; }
    MVTB_CUDA(cudaSetDevice(p->device));
    for (int i = 0; i < p->prof_n; ++i) {
        MVTB_CUDA(cudaEventSynchronize(p->prof_ev[2 * i + 1]));
        float ms = 0.f;
        MVTB_CUDA(cudaEventElapsedTime(&ms, p->prof_ev[2 * i], p->prof_ev[2 * i + 1]));
        p->prof_ms[p->prof_kind[i]] += ms;
        p->prof_cnt[p->prof_kind[i]] += 1;
    }
    p->prof_n = 0;
    for (int k = 0; k < MVTB_K_KINDS; ++k) { ms_sum[k] = p->prof_ms[k]; counts[k] = p->prof_cnt[k]; }
    return MVTB_OK;
}

extern "C" size_t mvtb_plan_workspace_bytes(const mvtb_plan* p) { return p ? p->ws_bytes : 0; }
