"""Host side of the RNG contract (csrc/om_math.cuh ``om_draw``): Philox4x32-10, key = seed, counter =
(env_id, reset_count, stream, 0); integers by multiply-shift.  Plain Python integers -- used for the few draws the HOST
takes (which model a MultiMuJoCo reset switches to, loco_env_base.py:586-589)."""

_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF

STREAM_MODEL_RESET = 32


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = (int(x) & _MASK for x in counter)
    k0, k1 = (int(x) & _MASK for x in key)
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def philox_randint(seed, env_id, count, stream, n):
    """First output word of the contract's draw -> integer in [0, n)."""
    seed = int(seed)
    w = philox4x32_10((env_id, count, stream, 0), (seed & _MASK, (seed >> 32) & _MASK))
    return (w[0] * int(n)) >> 32
