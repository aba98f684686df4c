"""Oracle: learner-side pieces of the hot path in float64 NumPy (TEST INFRASTRUCTURE).

* ``ppo_returns``          restates ``PPOBuffer.finish_path`` (reference ``rl/algos/ppo.py:68-84``) with
  the bootstrap ``(not done) * V(s_T)`` of ``:195-196`` and the advantage normalisation of ``:334-336``
  (``torch.std`` -> unbiased, ``eps=1e-5``).  Pinned: reference class imported by ``tools/gen_golden.py``.
* ``compute_gae``          restates ``mushroom_rl.utils.value_functions.compute_gae`` (mushroom-rl>=1.10,
  not in tree), called at ``imitation_lib/imitation/gail_TRPO.py:126-128``
  followed by ``(adv-mean)/(np.std(adv)+1e-8)``.  PARITY UNPINNED by the reference (third party).
* ``Standardizer``         restates ``imitation_lib/utils/networks.py:48-81``.
* ``RunningMeanStd`` / ``normalization_params``  restate ``rl/envs/normalize.py:190-208`` and ``:48``.
* ``gail_disc_loss`` / ``vdb_kl`` / ``vdb_loss``  restate ``imitation_lib/utils/math.py:11-86`` (pinned: the reference
  classes, ``tests/golden/disc_loss_ref.npz``); ``expert_indices`` is the expert-minibatch RNG contract.
* ``ppo_losses``           restates ``PPO.update_policy`` (``rl/algos/ppo.py:231-282``); pinned: the reference method run
  with stand-in actor / critic modules (``tests/golden/ppo_loss_ref.npz``).
* ``vail_forward`` / ``gail_forward`` / ``discrim_reward``  restate ``networks.py:258-284``,
  ``:208-234`` and ``gail_TRPO.py:320-327`` for the shapes fixed by
  ``examples/imitation_learning/utils.py:151-179`` + ``confs.yaml:113-130``.
"""
import numpy as np


# ------------------------------------------------------------------ PPO returns (G1)
def ppo_returns(rewards, values, done_last, v_boot, gamma):
    """One path per env of fixed length T.  rewards/values [T,N]; done_last [N] bool (episode ended by
    termination at the last step); v_boot [N] = V(s_T).  Returns (returns [T,N], advantages [T,N])."""
    rewards = np.asarray(rewards, np.float64)
    T = rewards.shape[0]
    R = np.where(done_last, 0.0, 1.0) * np.asarray(v_boot, np.float64)      # ppo.py:196
    ret = np.empty_like(rewards)
    for t in range(T - 1, -1, -1):                                          # ppo.py:75-77
        R = gamma * R + rewards[t]
        ret[t] = R
    return ret, ret - np.asarray(values, np.float64)                        # ppo.py:335


def ppo_returns_segmented(rewards, values, done, v_next, gamma):
    """Fixed-horizon rollout buffer with episode ends inside it (``sample`` loop ppo.py:166-196):
    ``done[t]`` marks the last step of a path that terminated (bootstrap 0); a path cut by the end of
    the buffer bootstraps with ``v_next[T-1]`` (= V of the state after the last step)."""
    rewards = np.asarray(rewards, np.float64)
    T = rewards.shape[0]
    ret = np.empty_like(rewards)
    R = np.where(done[T - 1], 0.0, np.asarray(v_next, np.float64)[T - 1])
    for t in range(T - 1, -1, -1):
        if t < T - 1:
            R = np.where(done[t], 0.0, R)
        R = gamma * R + rewards[t]
        ret[t] = R
    return ret, ret - np.asarray(values, np.float64)


def normalize_advantage_ppo(adv, eps=1e-5):
    """ppo.py:336: torch.std is the unbiased estimator."""
    adv = np.asarray(adv, np.float64)
    return (adv - adv.mean()) / (adv.std(ddof=1) + eps)


# ------------------------------------------------------------------ GAE (G2)
def compute_gae(v, v_next, r, absorbing, last, gamma, lam):
    """mushroom_rl compute_gae over a flat dataset of n transitions (loop from n-1 down)."""
    v = np.asarray(v, np.float64)
    v_next = np.asarray(v_next, np.float64)
    r = np.asarray(r, np.float64)
    n = len(r)
    adv = np.empty(n)
    for k in range(n - 1, -1, -1):
        if last[k] or k == n - 1:
            adv[k] = r[k] - v[k]
            if not absorbing[k]:
                adv[k] += gamma * v_next[k]
        else:
            adv[k] = r[k] + gamma * v_next[k] - v[k] + gamma * lam * adv[k + 1]
    return adv + v, adv


def compute_gae_batched(v, v_next, r, absorbing, last, gamma, lam):
    """Same recurrence for a time-major rollout buffer [T,N] of N independent envs; the buffer end
    (t == T-1) acts as the ``k == n-1`` case for every env."""
    v, v_next, r = (np.asarray(a, np.float64) for a in (v, v_next, r))
    T = r.shape[0]
    adv = np.empty_like(r)
    for t in range(T - 1, -1, -1):
        term = last[t] if t < T - 1 else np.ones_like(last[t], dtype=bool)
        boot = np.where(absorbing[t], 0.0, gamma * v_next[t])
        tail = r[t] - v[t] + boot
        if t < T - 1:
            cont = r[t] + gamma * v_next[t] - v[t] + gamma * lam * adv[t + 1]
            adv[t] = np.where(term, tail, cont)
        else:
            adv[t] = tail
    return adv + v, adv


def normalize_advantage_gail(adv):
    """gail_TRPO.py:128: np.std is the population estimator."""
    adv = np.asarray(adv, np.float64)
    return (adv - adv.mean()) / (adv.std() + 1e-8)


# ------------------------------------------------------------------ statistics (S1)
class Standardizer:
    """networks.py:48-81 (running sums start at 0, 1e-2, 1e-2; std floor sqrt(1e-2))."""

    def __init__(self):
        self._sum, self._sumsq, self._count = 0.0, 1e-2, 1e-2
        self.mean, self.std = 0.0, 1.0

    def update_mean_std(self, x):
        x = np.asarray(x, np.float64)
        self._sum = self._sum + x.sum(axis=0).ravel()
        self._sumsq = self._sumsq + np.square(x).sum(axis=0).ravel()
        self._count = self._count + np.array([len(x)])
        self.mean = self._sum / self._count
        self.std = np.sqrt(np.maximum(self._sumsq / self._count - np.square(self.mean), 1e-2))

    def forward(self, x):
        self.update_mean_std(x)
        return (np.asarray(x, np.float64) - self.mean) / self.std


class RunningMeanStd:
    """rl/envs/normalize.py:190-208 (parallel-variance merge)."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.zeros(shape, "float64")
        self.count = epsilon

    def update(self, x):
        bm, bv, bc = np.mean(x, axis=0), np.var(x, axis=0), x.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        M2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.mean, self.var, self.count = new_mean, M2 / tot, tot


def normalization_params(states):
    """rl/envs/normalize.py:48."""
    return np.mean(states, axis=0), np.sqrt(np.var(states, axis=0) + 1e-8)


# ------------------------------------------------------------------ discriminator reward (D1)
def _linear(x, w, b):
    return x @ np.asarray(w, np.float64).T + np.asarray(b, np.float64)


def vail_forward(p, s, eps, mean, std):
    """VariationalNet.forward (networks.py:258-284) for the H1 VAIL shapes, states only.
    p: dict w1[256,32] b1 w2[128,256] b2 wmu[128,128] bmu wlv[128,128] blv wd[1,128] bd."""
    x = (np.asarray(s, np.float64) - mean) / std                       # Standardizer.forward :73-74
    h = np.maximum(_linear(x, p["w1"], p["b1"]), 0.0)                  # encoder relu
    h = np.maximum(_linear(h, p["w2"], p["b2"]), 0.0)                  # encoder output relu
    mu = _linear(h, p["wmu"], p["bmu"])
    logvar = _linear(h, p["wlv"], p["blv"])
    z = mu + np.exp(logvar / 2) * np.asarray(eps, np.float64)          # reparameterize :21-24
    d = _linear(z, p["wd"], p["bd"])                                   # decoder, identity
    return d[..., 0], mu, logvar


def gail_forward(p, s, mean, std):
    """DiscriminatorNetwork.forward (networks.py:208-234), tanh-tanh-identity, 32-512-256-1."""
    x = (np.asarray(s, np.float64) - mean) / std
    h = np.tanh(_linear(x, p["w1"], p["b1"]))
    h = np.tanh(_linear(h, p["w2"], p["b2"]))
    return _linear(h, p["w3"], p["b3"])[..., 0]


def discrim_reward(d):
    """gail_TRPO.py:326-327."""
    plcy_prob = 1.0 / (1.0 + np.exp(-np.asarray(d, np.float64)))
    return (-np.log(1.0 - plcy_prob + 1e-8)).astype(np.float32)


# ------------------------------------------------------------------ discriminator fit (N2)
def logit_bernoulli_entropy(x):
    """imitation_lib/utils/math.py:33-38: (1 - sigmoid x) x - logsigmoid x (stable form)."""
    x = np.asarray(x, np.float64)
    sp = np.log1p(np.exp(-np.abs(x)))
    sig = np.where(x >= 0, 1.0 / (1.0 + np.exp(-np.abs(x))), np.exp(-np.abs(x)) / (1.0 + np.exp(-np.abs(x))))
    return (1.0 - sig) * x - (np.minimum(x, 0.0) - sp)


def bce_with_logits(x, t):
    """math.py:24-26: max(x, 0) - x t + log(1 + exp(-|x|)), per sample."""
    x, t = np.asarray(x, np.float64), np.asarray(t, np.float64)
    return np.maximum(x, 0.0) - x * t + np.log1p(np.exp(-np.abs(x)))


def gail_disc_loss(logits, targets, entcoeff=1e-3):
    """GailDiscriminatorLoss.forward (math.py:22-31) and its gradient with respect to the logits."""
    x, t = np.asarray(logits, np.float64).ravel(), np.asarray(targets, np.float64).ravel()
    loss = bce_with_logits(x, t).mean() - entcoeff * logit_bernoulli_entropy(x).mean()
    sig = 1.0 / (1.0 + np.exp(-x))
    grad = ((sig - t) + entcoeff * x * sig * (1.0 - sig)) / x.size
    return loss, grad


def vdb_kl(mu, logvar):
    """VDBLoss.kl_divergence (math.py:83-86), per sample."""
    mu, logvar = np.asarray(mu, np.float64), np.asarray(logvar, np.float64)
    return 0.5 * np.sum(mu ** 2 + np.exp(logvar) - logvar - 1.0, axis=1)


def vdb_loss(logits, kl, targets, beta, info_constraint, lr_beta):
    """VDBLoss.forward + _update_beta (math.py:55-81, use_bernoulli_ent=False) -> (loss, new beta)."""
    x, t = np.asarray(logits, np.float64).ravel(), np.asarray(targets, np.float64).ravel()
    bottleneck = np.mean(kl) - info_constraint
    loss = bce_with_logits(x, t).mean() + beta * bottleneck
    return loss, max(0.0, beta + lr_beta * bottleneck)


STREAM_EXPERT = 48


def _mix32(x):
    x = np.asarray(x, np.uint64) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13); x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def expert_indices(seed, draw, batch, n_src):
    """The expert-minibatch contract (stands in for mushroom_rl's minibatch_generator, which shuffles with NumPy's
    global stream: gail_TRPO.py:175-206 takes the first ``batch`` rows of a fresh permutation).  Sample b reads row
    pi_e(b mod n_src), e = b // n_src; pi_e = 4-round balanced Feistel network over 2*half bits with cycle walking,
    round function mix32(r ^ key_i) (murmur3 finaliser), round keys = the four Philox words of
    (seed; counter = (e, draw, STREAM_EXPERT, 0))."""
    from . import philox
    bits = 2
    while (1 << bits) < n_src:
        bits += 1
    half = (bits + 1) // 2
    mask = np.uint64((1 << half) - 1)
    b = np.arange(batch, dtype=np.int64)
    epoch, v = b // n_src, (b % n_src).astype(np.uint64)
    keys = [k.astype(np.uint64) for k in philox.draw(seed, epoch, draw, STREAM_EXPERT)]
    todo = np.ones(batch, bool)
    while todo.any():
        l, r = v[todo] >> np.uint64(half), v[todo] & mask
        for k in keys:
            f = _mix32(r ^ k[todo]) & mask
            l, r = r, l ^ f
        v[todo] = (l << np.uint64(half)) | r
        todo = v >= np.uint64(n_src)
    return v.astype(np.int64)


# ------------------------------------------------------------------ PPO minibatch losses (N3)
def ppo_losses(logp, old_logp, adv, mask, values, returns, entropy, det_actions, mirror_raw, mirrored_acts, clip, vf_coeff):
    """PPO.update_policy (rl/algos/ppo.py:231-282) from the quantities the networks produced: clipped surrogate, entropy
    penalty, value loss, approximate KL, mirror-symmetry loss (mirror_action applied to policy(mirror(obs)),
    wrappers.py:54-55) and clip fraction; plus d(actor_loss)/d(log_probs) and d(critic_loss)/d(values)."""
    f = lambda a: np.asarray(a, np.float64)
    logp, old_logp, adv, mask, values, returns = (f(a).reshape(-1) for a in (logp, old_logp, adv, mask, values, returns))
    n = logp.size
    ratio = np.exp(logp - old_logp)
    A = adv * mask
    cpi, clp = ratio * A, np.clip(ratio, 1 - clip, 1 + clip) * A
    out = dict(actor_loss=-np.minimum(cpi, clp).mean(),
               entropy_penalty=-(f(entropy) * mask[:, None]).mean(),
               critic_loss=vf_coeff * np.mean((returns - values) ** 2),
               approx_kl=np.mean((ratio - 1) - (logp - old_logp)),
               clip_fraction=np.mean(np.abs(ratio - 1) > clip))
    mirrored = mirror(f(mirror_raw), mirrored_acts)
    out["mirror_loss"] = np.mean((f(det_actions) - mirrored) ** 2)
    inside = (ratio >= 1 - clip) & (ratio <= 1 + clip)
    out["dlogp"] = -np.where(inside | (cpi < clp), ratio * A, 0.0) / n
    out["dvalues"] = vf_coeff * 2.0 * (values - returns) / n
    return out


# ------------------------------------------------------------------ mirror symmetry (N3)
def symmetry_matrix(mirrored):
    """rl/envs/wrappers.py:75-82."""
    numel = len(mirrored)
    mat = np.zeros((numel, numel))
    for i, j in zip(np.arange(numel), np.abs(np.array(mirrored).astype(int))):
        mat[i, j] = np.sign(mirrored[i])
    return mat


def mirror(x, mirrored, clock_inds=()):
    """SymmetricEnv.mirror_observation / mirror_action (wrappers.py:51-55) and, with ``clock_inds``,
    mirror_clock_observation (:60-72)."""
    y = np.asarray(x, np.float64) @ symmetry_matrix(mirrored)
    for c in clock_inds:
        y[:, c] = np.sin(np.arcsin(y[:, c]) + np.pi)
    return y
