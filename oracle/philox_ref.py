"""TEST INFRASTRUCTURE — host restatement of the noise-injection extension's generator.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123), the
4x32 variant cuRAND also uses — NOT numpy's ``np.random.Philox`` (4x64).  Pinned by the Random123 known-answer
vectors (tests/test_oracle.py).  The reference has no noise injection at all: parity unpinned by the
reference; this oracle pins the bit stream, Box-Muller values are compared within fp32 tolerance.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
  """ctr: (..., 4) uint32-valued array, key: (k0, k1).  Returns (..., 4) uint64 array of 32-bit words."""
  c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
  k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
  for _ in range(10):
    p0, p1 = M0 * c[0], M1 * c[2]
    hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
    c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
    k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
  return np.stack(c, axis=-1)


def words(seed, offset, nquads):
  q = np.arange(nquads, dtype=np.uint64)
  ctr = np.stack([q & MASK, q >> np.uint64(32), np.full(nquads, offset & 0xFFFFFFFF, np.uint64),
                  np.full(nquads, (offset >> 32) & 0xFFFFFFFF, np.uint64)], axis=-1)
  return philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def normals(seed, offset, n):
  """n standard normals in the device's element order (fp32 Box-Muller on word pairs (0,1) and (2,3))."""
  w = words(seed, offset, (n + 3) // 4).astype(np.float32)
  u = (w + np.float32(0.5)) * np.float32(2.0 ** -32)
  r0, r1 = np.sqrt(np.float32(-2) * np.log(u[:, 0])), np.sqrt(np.float32(-2) * np.log(u[:, 2]))
  t0, t1 = np.float32(2 * np.pi) * u[:, 1], np.float32(2 * np.pi) * u[:, 3]
  z = np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1).astype(np.float32)
  return z.reshape(-1)[:n]
