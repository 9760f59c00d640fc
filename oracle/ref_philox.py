"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the library's counter-based
noise generator (stag_b200/csrc/noise.cuh), used by the GPU tests to check the
fused RNG path bit-for-bit on the integer stream and to ~1e-5 on the transformed
variates.  This part has no counterpart in the reference (which calls torch's
global Philox generator through torch.distributions: stag/layers.py:117-127,
torch/distributions/normal.py:82-85, uniform.py:85-88, bernoulli.py:116-119);
bitwise parity with torch's stream is not a goal (SURVEY.md 8(c) "RNG") -- the law
is what must agree, and tests/test_gpu_rng.py checks the law statistically.

Generator: Philox4x32 with ROUNDS = 7 rounds (Salmon et al., SC'11; Random123 constants).
The round function is pinned against the Random123 known-answer vectors at 10 rounds
(tests/test_oracle_cpu.py); 7 is the smallest Crush-resistant round count (ibid., table 2).

Counter / key layout (one 128-bit block b = 8 channels of one edge):
    ctr = (b, eid, sample, offset_lo)      key = (seed_lo, seed_hi ^ offset_hi)
    channels of block b: c(b)+{0..3} and c(b)+32+{0..3}, c(b) = 64*(b//8) + 4*(b%8)
Each output word r_i (i = 0..3) gives slot 2i (from its low 16 bits h_lo) and slot 2i+1 (from its
high 16 bits h_hi); slot j < 4 is channel c(b)+j, slot j >= 4 is channel c(b)+32+(j-4):
    uniform  : u = h / 65536                                          in [0, 1)
    normal   : u1 = (h_lo + 1/2) / 65536,  rad = sqrt(float32(-2 ln 2) * log2(u1))
               ang = fma(float32(2^23 + h_hi), float32(2 pi / 65536), float32(-2 pi (128 - 2^-17)))
                   ~ 2 pi (h_hi + 1/2) / 65536  (the fp32 fma is part of the definition)
               z(slot 2i) = rad * cos(ang),  z(slot 2i+1) = rad * sin(ang)
    bernoulli: 1 if u < p else 0
A per-edge noise (K == 1) uses q = 0 and the first variate only.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


ROUNDS = 7


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=ROUNDS):
    """Vectorised Philox4x32-<rounds>.  All arguments are broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    return philox4x32(c0, c1, c2, c3, k0, k1, rounds=10)


def raw_block(eid, q, sample, seed, offset):
    """The four 32-bit words for (edge, channel-quad, sample) under (seed, offset)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    offset = int(offset) & 0xFFFFFFFFFFFFFFFF
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) ^ (offset >> 32)
    return philox4x32(q, eid, sample, np.uint32(offset & 0xFFFFFFFF), k0, k1)


def n_blocks(K):
    rem = K % 64
    return 8 * (K // 64) + min(8, (rem + 3) // 4)


def block_channels(nblk):
    """[nblk, 8] channel served by each slot of each block."""
    b = np.arange(nblk)
    c = 64 * (b // 8) + 4 * (b % 8)
    return np.concatenate([c[:, None] + np.arange(4), c[:, None] + 32 + np.arange(4)], axis=1)


def _to_channels(slots, K):
    """[E, nblk, 8] per-slot values -> [E, K] per-channel values."""
    E, nblk, _ = slots.shape
    ch = block_channels(nblk)
    out = np.zeros((E, max(int(ch.max()) + 1, K)), dtype=slots.dtype)
    out[:, ch.reshape(-1)] = slots.reshape(E, -1)
    return out[:, :K]


def _edge_ids(num_edges):
    """`num_edges` is either a count (edges 0 .. E-1, original order) or an array of edge ids (the variates of
    exactly those edges: row subsets of graphs too large to materialise [E,K] for)."""
    if np.ndim(num_edges) == 0:
        return np.arange(int(num_edges), dtype=np.uint32)
    return np.asarray(num_edges).astype(np.uint32).reshape(-1)


def halves(num_edges, K, sample, seed, offset):
    """uint32 [E, nblk, 8] -- the 16-bit integer h of every slot of every block (ORIGINAL edge order)."""
    nblk = n_blocks(K)
    eid = _edge_ids(num_edges)[:, None]
    num_edges = eid.shape[0]
    q = np.arange(nblk, dtype=np.uint32)[None, :]
    r = np.stack(raw_block(eid, q, np.uint32(sample), seed, offset), axis=-1)      # [E, nblk, 4]
    h = np.stack([r & np.uint32(0xFFFF), r >> np.uint32(16)], axis=-1)               # [E, nblk, 4, 2]
    return h.reshape(num_edges, nblk, 8)


def uniform(num_edges, K, sample, seed, offset):
    h = halves(num_edges, K, sample, seed, offset).astype(np.float32) * np.float32(2.0 ** -16)
    return _to_channels(h, K)


K_ANG = np.float64(np.float32(2.0 * np.pi / 65536.0))
C_ANG = np.float64(np.float32(-2.0 * np.pi * (128.0 - 2.0 ** -17)))
M2LN2 = np.float64(np.float32(-2.0 * np.log(2.0)))


def std_normal(num_edges, K, sample, seed, offset):
    nblk = n_blocks(K)
    h = halves(num_edges, K, sample, seed, offset)
    num_edges = h.shape[0]
    h = h.reshape(num_edges, nblk * 4, 2).astype(np.float64)
    u1 = (h[..., 0] + 0.5) * 2.0 ** -16
    rad = np.sqrt(M2LN2 * np.log2(u1))
    ang = np.float32((8388608.0 + h[..., 1]) * K_ANG + C_ANG).astype(np.float64)   # exact fp32 fma
    z = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=-1).reshape(num_edges, nblk, 8)
    return _to_channels(z.astype(np.float32), K)


# ---- STAG_NOISE_NORMAL_HADAMARD (stag_b200/csrc/spmm_tc.cuh, noise.cuh wh_*) ---------------------------------
# byte k (0..127) of an (edge, sample, 128-channel group g) = byte k % 16 (little endian) of Philox block
# 8 g + k // 16; FP8 e4m3 code = (byte & 0xCD) | 0x12; z[c] = WH_INV_SD * sum_k (-1)^popcount(c & k) value(code_k).
WH_AND, WH_OR = 0xCD, 0x12
WH_INV_SD = np.float32(0.006240209594902129)   # 1 / sqrt(128 * 105186885 / 524288)


def e4m3_value(code):
    """Value of an e4m3 code (normal numbers; the masked codes never are subnormal or NaN)."""
    code = np.asarray(code, dtype=np.int64)
    sign = np.where(code & 0x80, -1.0, 1.0)
    e = (code >> 3) & 0xF
    m = code & 7
    return sign * (1.0 + m / 8.0) * np.exp2(e - 7.0)


def hadamard_matrix(n=128):
    k = np.arange(n)
    par = np.zeros((n, n), dtype=np.int64)
    x = k[:, None] & k[None, :]
    while x.any():
        par ^= x & 1
        x >>= 1
    return 1.0 - 2.0 * par


def hadamard_sums(num_edges, K, sample, seed, offset):
    """float64 [E, K] raw sums (exact: multiples of 2^-8 below 2^12), original edge order."""
    assert K % 128 == 0
    G = K // 128
    eid = _edge_ids(num_edges)[:, None]        # a count, or an array of edge ids
    num_edges = eid.shape[0]
    blk = np.arange(8 * G, dtype=np.uint32)[None, :]
    r = np.stack(raw_block(eid, blk, np.uint32(sample), seed, offset), axis=-1)          # [E, 8G, 4] words
    by = np.stack([(r >> np.uint32(8 * b)) & np.uint32(0xFF) for b in range(4)], axis=-1)  # [E, 8G, 4, 4]
    code = (by.reshape(num_edges, G, 128).astype(np.int64) & WH_AND) | WH_OR
    v = e4m3_value(code)                                                                  # [E, G, 128]
    return (v @ hadamard_matrix().T).reshape(num_edges, K)


def hadamard_normal(num_edges, K, sample, seed, offset):
    return (hadamard_sums(num_edges, K, sample, seed, offset).astype(np.float32) * WH_INV_SD).astype(np.float32)


def noise(kind, num_edges, K, sample, seed, offset, p0=None, p1=None, relu=False):
    """w [E,K] float32 for a distribution kind in {'normal','uniform','bernoulli'};
    p0/p1 broadcastable to [E,K] (loc/scale, low/high, probs)."""
    if kind == "normal":
        w = np.float32(p0) + std_normal(num_edges, K, sample, seed, offset) * np.float32(p1)
    elif kind == "normal_hadamard":
        # the kernels evaluate fma(sum, scale * WH_INV_SD, loc) in fp32
        b = (np.asarray(p1, dtype=np.float32) * WH_INV_SD).astype(np.float32)
        w = (hadamard_sums(num_edges, K, sample, seed, offset) * b.astype(np.float64)
             + np.asarray(p0, dtype=np.float32).astype(np.float64)).astype(np.float32)
    elif kind == "uniform":
        w = np.float32(p0) + uniform(num_edges, K, sample, seed, offset) * (np.float32(p1) - np.float32(p0))
    elif kind == "bernoulli":
        w = (uniform(num_edges, K, sample, seed, offset) < np.float32(p0)).astype(np.float32)
    else:
        raise KeyError(kind)
    w = np.asarray(w, dtype=np.float32)
    return np.maximum(w, 0) if relu else w


# Random123 known-answer vectors for philox4x32-10 (kat_vectors): (ctr, key, expected)
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
