"""numpy restatement of tensorbnn_b200/csrc/philox.cuh (Philox4x32-10 + Box-Muller),
used to check the on-device momentum / uniform streams bit-for-bit at the integer level
and to float tolerance after Box-Muller."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_MAIN, STREAM_HYPER = 0x0, 0x48595045
UNIFORM_BLOCK = 0xFFFFFFFF
MASK = 0xFFFFFFFF


def philox4x32_10(c, k0, k1):
    x, y, z, w = [int(v) & MASK for v in c]
    k0 &= MASK
    k1 &= MASK
    for _ in range(10):
        p0, p1 = M0 * x, M1 * z
        x, y, z, w = ((p1 >> 32) ^ y ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ w ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return x, y, z, w


def u01f(x):
    return (np.float32(x >> 8) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def u01d(hi, lo):
    v = ((hi << 32) | lo) >> 11
    return (float(v) + 0.5) / 9007199254740992.0


def normals_f32(seed, tag, call, chain, n):
    out = np.empty(n, dtype=np.float64)
    for j in range((n + 3) // 4):
        r = philox4x32_10((j, chain, call & MASK, call >> 32), seed & MASK, ((seed >> 32) & MASK) ^ tag)
        u = [float(u01f(v)) for v in r]
        r0, r1 = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
        vals = [r0 * np.cos(2 * np.pi * u[1]), r0 * np.sin(2 * np.pi * u[1]),
                r1 * np.cos(2 * np.pi * u[3]), r1 * np.sin(2 * np.pi * u[3])]
        for q in range(4):
            if 4 * j + q < n:
                out[4 * j + q] = vals[q]
    return out


def normals_f64(seed, tag, call, chain, n):
    out = np.empty(n, dtype=np.float64)
    for j in range((n + 1) // 2):
        r = philox4x32_10((j, chain, call & MASK, call >> 32), seed & MASK, ((seed >> 32) & MASK) ^ tag)
        u0, u1 = u01d(r[0], r[1]), u01d(r[2], r[3])
        rr = np.sqrt(-2.0 * np.log(u0))
        vals = [rr * np.cos(2 * np.pi * u1), rr * np.sin(2 * np.pi * u1)]
        for q in range(2):
            if 2 * j + q < n:
                out[2 * j + q] = vals[q]
    return out


def uniform(seed, tag, call, chain):
    r = philox4x32_10((UNIFORM_BLOCK, chain, call & MASK, call >> 32), seed & MASK,
                      ((seed >> 32) & MASK) ^ tag)
    return u01d(r[0], r[1])
