// sp_sampler.cuh — the geometric-gap salt-and-pepper sampler, one warp per span of MVTB_SP_SPAN consecutive voxels.
//
// Hits are i.i.d. Bernoulli(p) (what `u <= p` gives the reference, F:478-479) drawn at a cost proportional to p: the
// number of untouched voxels before the next hit is geometric, P(gap >= k) = (1-p)^k, obtained by inverse CDF from a
// 32-bit Philox word w against the integer table T[k] = floor(2^32 (1 - (1-p)^(k+1))), k < 256 (exact integer
// compares; a word with w >= T[255] means "no hit in the next 256 voxels", after which the walk goes on afresh --
// the geometric law has no memory).  One more random bit per hit picks salt or pepper.
//
// A span is ONE stream consumed by the 32 lanes of a warp in lock step: in iteration i lane l runs
// Philox4x32-10(counter = (span id lo, span id hi, 32 i + l, 0x5351), key = seed), turns words x, y, z into three
// (advance, hit) pairs (coins: bits 0..2 of word w), and a warp prefix sum places them behind everything lanes
// 0..l-1 and the earlier iterations produced.  Every lane executes the same instructions until the span is covered
// (no divergent trip counts: the round-1 sampler, one 256-voxel block per thread, executed 3.3 M warp instructions
// per 240x240x155 volume at p = 0.05, this one ~0.9 M), and nothing goes through shared-memory lists.
// oracle/philox_ref.py:sparse_hits restates this bit for bit in numpy.
#pragma once
#include "philox.cuh"

#define MVTB_SP_TAG 0x5351u

namespace mvtb {

// smallest k with w < T[k] (T non-decreasing), 256 if none: one MUFU.LG2 guess settled by the exact table
__device__ __forceinline__ int sp_invert(unsigned w, const unsigned* __restrict__ sT, float inv_log2q) {
    const float v = (float)(~w) * 2.3283064365386963e-10f;          // 1 - w / 2^32 without cancellation
#ifdef MVTB_EMU
    const float kf = log2f(v) * inv_log2q;
#else
    const float kf = __log2f(v) * inv_log2q;
#endif
    int k = (int)fminf(fmaxf(kf, 0.f), (float)MVTB_SP_BLOCK);       // NaN -> 0 (fmaxf) -> corrected below
    while (k > 0 && w < sT[k - 1]) --k;
    while (k < MVTB_SP_BLOCK && w >= sT[k]) ++k;
    return k;
}

// NS consecutive spans (span0 .. span0 + NS - 1 of the sample, those < n_spans) walked by one warp in lock step: the
// per-span chains (Philox -> inversion -> prefix sum -> stores) are independent, so NS of them overlap their latencies.
// xs: first voxel of the sample; gs0: global id (Philox counter) of span0.  The result does not depend on NS.
// stage: null (hits are stored as they are found: the fused kernel, whose output lines are in L2), or this warp's
// kSpStageIters * 96 shared-memory slots: hits are parked there in stream order and written out 32 consecutive hits
// per store instruction (~2.5 KB of the volume at p = 0.05), which keeps the read-modify-writes of a DRAM page together
// -- stored straight from the walk, a warp's 32 stores land 250 voxels apart each and the select pass, alone on HBM,
// takes 10.7 instead of 6.6 us per volume.
static const int kSpStageIters = 10;
template <int NS>
__device__ __forceinline__ void sp_walk_spans(float* __restrict__ xs, unsigned long long n_per_sample, unsigned span0,
                                              unsigned n_spans, unsigned long long gs0, uint2 key,
                                              const unsigned* __restrict__ sT, float inv_log2q, float lo, float hi, int lane,
                                              unsigned short* __restrict__ stage = nullptr) {
    int len[NS], base[NS];
    float* xp[NS];
    bool live = false;
    MVTB_UNROLL
    for (int u = 0; u < NS; ++u) {
        const unsigned long long j0 = (unsigned long long)(span0 + u) * MVTB_SP_SPAN;
        len[u] = 0;
        if (span0 + u < n_spans)
            len[u] = (int)((n_per_sample - j0) < (unsigned long long)MVTB_SP_SPAN ? (n_per_sample - j0) : (unsigned long long)MVTB_SP_SPAN);
        xp[u] = xs + j0;
        base[u] = 0;                                                 // voxels of the span already walked
        live = live || len[u] > 0;
    }
    unsigned iter = 0;
    int staged = 0;                                                  // iterations parked in `stage` (NS == 1 only)
    while (live) {
        // Straight-line phases over all NS chains so that ptxas interleaves them (a warp issues in order: independent
        // chains only overlap if their instructions alternate).  A span that is already covered keeps computing; its
        // stores are all past its end.
        uint4 r[NS];
        MVTB_UNROLL
        for (int u = 0; u < NS; ++u) {
            const unsigned long long gs = gs0 + (unsigned long long)u;
            r[u] = Philox::run(make_uint4((unsigned)gs, (unsigned)(gs >> 32), iter * 32u + (unsigned)lane, MVTB_SP_TAG), key);
        }
        int pre[NS][3], run[NS];                                     // inclusive prefix of this lane's advances
        unsigned hitm[NS];
        MVTB_UNROLL
        for (int u = 0; u < NS; ++u) {
            const unsigned words[3] = {r[u].x, r[u].y, r[u].z};
            run[u] = 0;
            hitm[u] = 0;
            MVTB_UNROLL
            for (int j = 0; j < 3; ++j) {
                const int k = sp_invert(words[j], sT, inv_log2q);
                const bool h = k < MVTB_SP_BLOCK;
                hitm[u] |= h ? (1u << j) : 0u;
                run[u] += h ? k + 1 : MVTB_SP_BLOCK;
                pre[u][j] = run[u];
            }
        }
        int incl[NS];                                                // warp inclusive scans of the lane totals
        MVTB_UNROLL
        for (int u = 0; u < NS; ++u) incl[u] = run[u];
        MVTB_UNROLL
        for (int o = 1; o < 32; o <<= 1) {
            MVTB_UNROLL
            for (int u = 0; u < NS; ++u) {
                const int t = __shfl_up_sync(0xffffffffu, incl[u], o);
                if (lane >= o) incl[u] += t;
            }
        }
        live = false;
        MVTB_UNROLL
        for (int u = 0; u < NS; ++u) {
            const int start = base[u] + incl[u] - run[u];            // voxels walked before this lane's first word
            MVTB_UNROLL
            for (int j = 0; j < 3; ++j) {
                const int pos = start + pre[u][j] - 1;
                const bool h = ((hitm[u] >> j) & 1u) && pos < len[u];
                if (stage == nullptr) {
                    if (h) xp[u][pos] = ((r[u].w >> j) & 1u) ? hi : lo;
                } else {                                             // position (13 bits) | coin << 13, 0xffff = nothing
                    stage[staged * 96 + lane * 3 + j] = h ? (unsigned short)(pos | (((r[u].w >> j) & 1u) << 13)) : (unsigned short)0xffffu;
                }
            }
            base[u] += __shfl_sync(0xffffffffu, incl[u], 31);
            live = live || base[u] < len[u];
        }
        ++iter;
        if (stage != nullptr) {
            ++staged;
            if (staged == kSpStageIters || !live) {
                __syncwarp();
                for (int i = lane; i < staged * 96; i += 32) {
                    const unsigned e = stage[i];
                    if (e != 0xffffu) xp[0][e & 0x1fffu] = (e >> 13) ? hi : lo;
                }
                __syncwarp();
                staged = 0;
            }
        }
    }
}

// One span of `len` voxels, global id `gs`, walked by a warp in lock step: the same stream as sp_walk_spans (same
// counters, words, coins and order), each hit handed to on_hit(position within the span, coin) instead of stored.
template <typename F>
__device__ __forceinline__ void sp_walk_span_cb(int len, unsigned long long gs, uint2 key, const unsigned* __restrict__ sT,
                                                float inv_log2q, int lane, F on_hit) {
    int base = 0;
    unsigned iter = 0;
    while (base < len) {
        const uint4 r = Philox::run(make_uint4((unsigned)gs, (unsigned)(gs >> 32), iter * 32u + (unsigned)lane, MVTB_SP_TAG), key);
        const unsigned words[3] = {r.x, r.y, r.z};
        int pre[3], run = 0;
        unsigned hitm = 0;
        MVTB_UNROLL
        for (int j = 0; j < 3; ++j) {
            const int k = sp_invert(words[j], sT, inv_log2q);
            const bool h = k < MVTB_SP_BLOCK;
            hitm |= h ? (1u << j) : 0u;
            run += h ? k + 1 : MVTB_SP_BLOCK;
            pre[j] = run;
        }
        int incl = run;
        MVTB_UNROLL
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int start = base + incl - run;
        MVTB_UNROLL
        for (int j = 0; j < 3; ++j) {
            const int pos = start + pre[j] - 1;
            if (((hitm >> j) & 1u) && pos < len) on_hit(pos, (r.w >> j) & 1u);
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
        ++iter;
    }
}

}  // namespace mvtb
