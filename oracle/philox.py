"""Oracle: counter-based RNG contract shared by CPU oracle and CUDA path (TEST INFRASTRUCTURE).

The reference draws reset indices from NumPy's global MT19937 stream
(``olympic_mujoco/utils/trajectory.py:304,311`` ``np.random.randint``); a batched implementation
cannot replay a global sequential stream, so the contract (hard part 4) is:

* generator: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11;
  Random123 known-answer vectors are checked in ``tests/test_oracle_philox.py``);
* key = (seed & 0xffffffff, seed >> 32); counter = (env_id, reset_count, stream, 0);
* integer in [0, n): ``(x * n) >> 32`` on one 32-bit output word (multiply-shift);
* float in [0,1): ``(x >> 8) * 2**-24``.

With one env and injected draws the semantics are the reference's (see oracle/trajectory.py).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_TRAJ_RESET = 0      # words: [traj_no, substep, -, -]
STREAM_A3_RESET = 16       # streams 16.. used by the A3 reset (see oracle/a3.py)


def philox4x32_10(ctr, key):
    """ctr: 4 arrays of uint32 (broadcastable), key: 2 arrays of uint32 -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in ctr]
    c = list(np.broadcast_arrays(*c))
    k0 = np.asarray(key[0], dtype=np.uint64) & MASK
    k1 = np.asarray(key[1], dtype=np.uint64) & MASK
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return [x.astype(np.uint32) for x in c]


def draw(seed, env_id, count, stream=0):
    """The contract's 4 output words for (seed, env, counter, stream)."""
    seed = int(seed)
    return philox4x32_10((env_id, count, stream, 0), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def to_int(x, n):
    """uint32 word -> integer in [0, n) by multiply-shift."""
    return ((x.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int32)


def to_unit(x):
    """uint32 word -> float64 in [0,1) with 24 bits (exactly representable in float32)."""
    return (x >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)
