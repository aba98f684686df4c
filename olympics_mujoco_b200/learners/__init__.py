"""Learner-side hooks of the hot path, named as the reference names them.

* ``compute_gae``            mushroom_rl.utils.value_functions.compute_gae, call site
  ``imitation_lib/imitation/gail_TRPO.py:126-128`` (+ the advantage normalisation of :128)
* ``finish_path`` / ``normalize_advantage_ppo``   ``PPOBuffer.finish_path`` ``rl/algos/ppo.py:68-84``, bootstrap :195-196,
  ``advantages = returns - values`` and normalisation :334-336
* ``DiscriminatorReward.make_discrim_reward``     ``GAIL.make_discrim_reward`` ``gail_TRPO.py:320-327`` with the
  ``Standardizer`` side effect of ``networks.py:68-81``

Everything operates on time-major ``[T, n]`` (or ``[n]``) CUDA tensors of this rank's env shard; the value network and
the policy stay the learner's (PyTorch)."""
from __future__ import annotations

import torch

from .. import distributed as D
from .. import kernels as Kn


def compute_gae(v, v_next, r, absorbing, last, gamma, lam, normalize=False):
    """-> (v_target, adv) like the reference; ``v`` / ``v_next`` are V(s), V(s') evaluated by the caller.  With
    ``normalize`` the advantage is standardised over ALL ranks' transitions: (adv - mean) / (np.std + 1e-8)."""
    v_target, adv = Kn.gae(r, v, v_next, absorbing, last, gamma, lam)
    if normalize:
        mom = D.all_reduce_moments(Kn.moments_scalar(adv))
        Kn.normalize(adv, Kn.adv_stats(mom, unbiased=False, eps=1e-8), out=adv)
    return v_target, adv


def finish_path(rewards, values, gamma, path_end=None, v_next=None, last_val=None):
    """-> (returns, advantages = returns - values) for a rollout buffer; ``last_val`` = (not done) * V(s_T)."""
    return Kn.ppo_returns(rewards, values, gamma, path_end=path_end, v_next=v_next, v_last=last_val)


def normalize_advantage_ppo(adv, eps=1e-5):
    """ppo.py:336 over every rank's buffer: (adv - mean) / (torch.std (unbiased) + eps), in place."""
    mom = D.all_reduce_moments(Kn.moments_scalar(adv))
    return Kn.normalize(adv, Kn.adv_stats(mom, unbiased=True, eps=eps), out=adv)


class DiscriminatorReward:
    """The reward side of GAIL / VAIL for the UnitreeH1 configuration (state-only discriminator,
    ``use_next_states=False, use_actions=False``; ``examples/imitation_learning/utils.py:151-179``)."""

    def __init__(self, kind, params, device="cuda", seed=0):
        self.kind = kind.lower()
        self.disc = Kn.Discriminator(self.kind, params)
        self.standardizer = D.Standardizer(self.disc.n_in, device=device)
        self._gen = torch.Generator(device=device).manual_seed(seed)

    def make_discrim_reward(self, state, action=None, next_state=None, eps=None, update_standardizer=True):
        """state [32, n] (SoA) -> reward [n] float32.  Like the reference, every call first folds the batch into the
        standardiser's running sums (all ranks), then standardises with the updated statistics."""
        if update_standardizer:
            self.standardizer.update(state)
        mean, std = self.standardizer.snapshot_f32()
        if self.kind == "vail" and eps is None:
            eps = torch.randn((self.disc.z, state.shape[-1]), device=state.device, generator=self._gen)
        return self.disc.reward(state, mean, std, eps=eps if self.kind == "vail" else None)
