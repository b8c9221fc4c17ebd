"""numpy Philox4x32-10 — TEST INFRASTRUCTURE ONLY.

Restates the published algorithm (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
as 1, 2, 3", SC'11; Random123 philox.h) and is pinned by Random123's own known-answer vectors
(kat_vectors, philox4x32 10 rounds), checked in tests/test_host_logic.py.  The CUDA kernel's
generator (csrc/voxel_ops.cu) is compared against this bit for bit.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)

# Random123 kat_vectors: (counter[4], key[2]) -> output[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(ctr, key):
    """ctr: (n,4) uint32, key: (n,2) uint32 -> (n,4) uint32."""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0 = key[:, 0].astype(np.uint32).copy()
    k1 = key[:, 1].astype(np.uint32).copy()
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1, hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
        with np.errstate(over="ignore"):
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([v.astype(np.uint32) for v in c], axis=1)


def uniform_f32(n, seed, offset):
    """The uniforms mvtb_salt_pepper_f32 / mvtb_philox_uniform_f32 produce for elements 0..n-1."""
    ng = (n + 3) // 4
    g = np.arange(ng, dtype=np.uint64) + np.uint64(offset)
    ctr = np.zeros((ng, 4), dtype=np.uint32)
    ctr[:, 0] = (g & MASK).astype(np.uint32)
    ctr[:, 1] = (g >> np.uint64(32)).astype(np.uint32)
    key = np.empty((ng, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key).reshape(-1)[:n]
    return ((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


# ----------------------------------------------------------------------------- sparse Bernoulli sampler
SP_BLOCK = 256


def sparse_table(p):
    """T[k] = floor(2^32 (1 - (1-p)^(k+1))) with the power by repeated float64 multiplication
    (restates mvtb_sparse_table, csrc/voxel_ops.cu)."""
    q = np.float64(1.0) - np.float64(np.float32(p))
    t = np.float64(1.0)
    out = np.zeros(SP_BLOCK, dtype=np.uint32)
    for k in range(SP_BLOCK):
        t = t * q
        v = np.floor((np.float64(1.0) - t) * np.float64(4294967296.0))
        out[k] = np.uint32(min(max(v, 0.0), 4294967295.0))
    return out


SP_SPAN = 8192
SP_TAG = 0x5351


def sparse_hits(n_per_sample, n_samples, seed, offset, p):
    """Positions (flat index into the whole buffer) and kinds (1 = salt/max, 0 = pepper/min) that
    mvtb_salt_pepper_sparse_f32 / mvtb_kspace_chain_sp_f32 touch (restates csrc/sp_sampler.cuh).

    Each span of SP_SPAN voxels is one stream: iteration i, lane l (0..31) takes Philox words x, y, z of counter
    (span id lo, span id hi, 32 i + l, SP_TAG); a word w gives k = #{T <= w}; k < 256: advance k + 1 and hit the voxel
    reached, else advance 256 without a hit; coins are bits 0..2 of word w.  Words are consumed in (i, l, x|y|z)
    order until the span is covered; hits past its end are dropped."""
    T = sparse_table(p).astype(np.uint64)
    sps = (n_per_sample + SP_SPAN - 1) // SP_SPAN
    key1 = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    pos_out, kind_out = [], []
    lanes = np.arange(32, dtype=np.uint32)
    for s in range(n_samples):
        for sp in range(sps):
            j0 = sp * SP_SPAN
            length = min(SP_SPAN, n_per_sample - j0)
            gs = offset + s * sps + sp
            base, it = 0, 0
            while base < length:
                ctr = np.zeros((32, 4), dtype=np.uint32)
                ctr[:, 0] = np.uint32(gs & 0xFFFFFFFF)
                ctr[:, 1] = np.uint32((gs >> 32) & 0xFFFFFFFF)
                ctr[:, 2] = np.uint32(it * 32) + lanes
                ctr[:, 3] = np.uint32(SP_TAG)
                r = philox4x32_10(ctr, np.broadcast_to(key1, (32, 2)))
                w = r[:, :3].astype(np.uint64).reshape(-1)                    # (lane, word) order
                k = np.searchsorted(T, w, side="right")                       # smallest k with w < T[k]; 256 if none
                hit = k < SP_BLOCK
                adv = np.where(hit, k + 1, SP_BLOCK)
                pos = base + np.cumsum(adv) - 1
                coin = ((r[:, 3][:, None] >> np.arange(3, dtype=np.uint32)[None, :]) & 1).reshape(-1)
                sel = hit & (pos < length)
                pos_out.append(s * n_per_sample + j0 + pos[sel])
                kind_out.append(coin[sel].astype(np.int64))
                base += int(adv.sum())
                it += 1
    if not pos_out:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    return np.concatenate(pos_out).astype(np.int64), np.concatenate(kind_out).astype(np.int64)
