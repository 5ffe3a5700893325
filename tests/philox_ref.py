"""numpy restatement of Philox4x32-10 (Salmon et al. 2011) and of the library's documented
counter mapping -- test infrastructure, used to reproduce the GPU samplers draw for draw."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
STREAM_BACKGROUND, STREAM_PROBSAMPLE, STREAM_BOOTSTRAP, STREAM_KNN = 0, 1, 2, 3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(x, np.uint64) & np.uint64(0xFFFFFFFF) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint64(k0 & 0xFFFFFFFF)
    k1 = np.uint64(k1 & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c


def u01_53(hi, lo):
    return ((hi >> np.uint64(5)).astype(np.float64) * 67108864.0 +
            (lo >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def seed_words(seed):
    return seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
