"""Synthetic Hardy-Weinberg genotypes in EIGENSOFT's 2-bit packed layout.

Counter-based generator keyed by (seed, snp, indiv): host (numpy, this file) and device
(csrc/synth.cu, `eb_synth_packed`) produce identical bytes, and any SNP shard can be generated
independently on its own GPU.

  key(s,i)  = seed*C0 + s*C1 + i*C2            (mod 2^64)
  h         = splitmix64_finalise(key)
  p_s       = 0.05 + 0.9 * (splitmix64_finalise(key(s, 2^32-1)) >> 11) / 2^53
  p_{s,k}   = clip(p_s + delta_k * (u_{s,k}-0.5)*sqrt(12) * sqrt(p_s(1-p_s)), 0.01, 0.99)   (pop k of indiv i)
  g         = (lo32(h) < T) + (hi32(h) < T),   T = uint32(p * 2^32)
  missing   = lo32(splitmix64_finalise(key ^ CM)) < uint32(rho * 2^32)  -> code 3

Packed layout (admutils.c:718-735, mcio.c:2788-2790 of the reference): SNP-major, rlen = max(48, ceil(N/4))
bytes per SNP, individual k in byte k>>2 at bits (3-(k&3))*2, codes 0/1/2 = allele count, 3 = missing.
"""
import os
import numpy as np

C0 = np.uint64(0x9E3779B97F4A7C15)
C1 = np.uint64(0xBF58476D1CE4E5B9)
C2 = np.uint64(0x94D049BB133111EB)
CM = np.uint64(0xD6E8FEB86659FD93)
CP = np.uint64(0xA0761D6478BD642F)


def _mix(z):
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30); z *= C1
    z ^= z >> np.uint64(27); z *= C2
    z ^= z >> np.uint64(31)
    return z


def rlen_for(numindivs):
    return max(48, (numindivs + 3) // 4)


def snp_freqs(seed, s0, nsnp):
    s = np.arange(s0, s0 + nsnp, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = np.uint64(seed) * C0 + s * C1 + np.uint64(0xFFFFFFFF) * C2
    return 0.05 + 0.9 * ((_mix(key) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0)


def pop_of(numindivs, npops):
    return (np.arange(numindivs, dtype=np.int64) * npops) // numindivs


def genotypes(seed, nsnp, numindivs, missing=0.0, npops=1, delta=0.0, s0=0, pop_delta=None):
    """int8 [nsnp, numindivs] with 0/1/2 and -1 for missing. `pop_delta` (len npops) overrides the common delta."""
    p = snp_freqs(seed, s0, nsnp)
    s = np.arange(s0, s0 + nsnp, dtype=np.uint64)[:, None]
    i = np.arange(numindivs, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        key = np.uint64(seed) * C0 + s * C1 + i * C2
    h = _mix(key)
    if npops > 1:
        d = np.full(npops, delta) if pop_delta is None else np.asarray(pop_delta, np.float64)
        k = np.arange(npops, dtype=np.uint64)[None, :]
        with np.errstate(over="ignore"):
            kk = (np.uint64(seed) * C0 + s * C1 + k * C2) ^ CP
        u = (_mix(kk) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        pk = p[:, None] + d[None, :] * ((u - 0.5) * np.sqrt(12.0)) * np.sqrt(p * (1 - p))[:, None]
        pk = np.minimum(np.maximum(pk, 0.01), 0.99)
        pi = pk[:, pop_of(numindivs, npops)]
    else:
        pi = np.broadcast_to(p[:, None], (nsnp, numindivs))
    T = (pi * 4294967296.0).astype(np.uint64)
    g = ((h & np.uint64(0xFFFFFFFF)) < T).astype(np.int8) + ((h >> np.uint64(32)) < T).astype(np.int8)
    if missing > 0:
        Tm = np.uint64(int(missing * 4294967296.0))
        hm = _mix(key ^ CM) & np.uint64(0xFFFFFFFF)
        g[hm < Tm] = -1
    return g


def pack(g):
    """int8 [nsnp, N] (-1 missing) -> uint8 [nsnp, rlen] in the reference's layout; pad genotypes are code 3."""
    nsnp, n = g.shape
    rl = rlen_for(n)
    codes = np.full((nsnp, rl * 4), 3, np.uint8)
    codes[:, :n] = np.where(g < 0, 3, g).astype(np.uint8)
    c = codes.reshape(nsnp, rl, 4)
    return ((c[:, :, 0] << 6) | (c[:, :, 1] << 4) | (c[:, :, 2] << 2) | c[:, :, 3]).astype(np.uint8)


def unpack(packed, numindivs):
    c = np.stack([(packed >> 6) & 3, (packed >> 4) & 3, (packed >> 2) & 3, packed & 3], axis=-1)
    c = c.reshape(packed.shape[0], -1)[:, :numindivs].astype(np.int8)
    c[c == 3] = -1
    return c


def packed_genotypes(seed, nsnp, numindivs, chunk=4096, **kw):
    rl = rlen_for(numindivs)
    out = np.empty((nsnp, rl), np.uint8)
    for a in range(0, nsnp, chunk):
        b = min(nsnp, a + chunk)
        out[a:b] = pack(genotypes(seed, b - a, numindivs, s0=a + kw.get("s0", 0), **{k: v for k, v in kw.items() if k != "s0"}))
    return out


def write_dataset(prefix, packed, numindivs, pops=None, npops=1):
    """Write PACKEDANCESTRYMAP .geno + .snp + .ind (mcio.c:2343-2440 layout; use `hashcheck: NO`)."""
    nsnp, rl = packed.shape
    assert rl == rlen_for(numindivs)
    hdr = bytearray(rl)
    s = ("GENO %7d %7d %x %x" % (numindivs, nsnp, 0, 0)).encode()
    hdr[:len(s)] = s
    with open(prefix + ".geno", "wb") as f:
        f.write(bytes(hdr)); f.write(packed.tobytes())
    with open(prefix + ".snp", "w") as f:
        per = max(1, (nsnp + 21) // 22)
        for k in range(nsnp):
            ch = k // per + 1; pos = (k % per) * 1000 + 1000
            f.write("%20s %2d %12.6f %12d A C\n" % ("rs%d" % k, ch, pos * 1e-8, pos))
    if pops is None:
        pops = ["Pop%d" % k for k in pop_of(numindivs, npops)]
    with open(prefix + ".ind", "w") as f:
        for k in range(numindivs):
            f.write("%20s U %s\n" % ("ind%d" % k, pops[k]))
    return prefix
