"""N > 1 host logic on CPU: two gloo ranks shard the envs, all-reduce their float64 moment partial sums and must
reproduce the single-process statistics of the reference formulas; the Philox contract makes the sharded resets the
same draws as the single-process ones."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import learner as L
from oracle import philox


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _moments(x):
    x = torch.as_tensor(x, dtype=torch.float64)
    return torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([float(x.shape[0])], dtype=torch.float64)])


def _worker(rank, world, port, n_total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from olympics_mujoco_b200 import distributed as D
    r, w, _ = D.init(backend="gloo")
    assert (r, w) == (rank, world) and D.world() == world
    env_id0, n_local = D.env_shard(n_total, rank, world)
    rng = np.random.default_rng(0)
    obs = rng.normal(1.0, 2.0, (n_total, 32))                  # every rank generates the global batch, keeps its shard
    adv = rng.normal(0.3, 1.5, (n_total, 1))
    mine = slice(env_id0, env_id0 + n_local)
    st = D.Standardizer(32, device="cpu")
    for _ in range(2):                                         # two fits: running sums accumulate
        mean, std = st.update_from_moments(_moments(obs[mine]))
    am = D.all_reduce_moments(_moments(adv[mine]))
    pm, ps = D.mean_std_from_moments(D.all_reduce_moments(_moments(obs[mine])), "ppo_obs")
    a_ppo = D.mean_std_from_moments(am, "adv_ppo")
    a_gail = D.mean_std_from_moments(am, "adv_gail")
    # sharded reset draws: the trajectory-reset words of this rank's envs under the Philox contract
    draws = np.stack(philox.draw(1234, np.arange(env_id0, env_id0 + n_local, dtype=np.uint32), np.uint32(0)), axis=1)
    q.put((rank, env_id0, n_local, mean.numpy(), std.numpy(), pm.numpy(), ps.numpy(),
           [float(x) for x in a_ppo], [float(x) for x in a_gail], draws))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1000, 1001])
def test_two_rank_moment_allreduce_matches_single_process(n_total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    obs = rng.normal(1.0, 2.0, (n_total, 32))
    adv = rng.normal(0.3, 1.5, (n_total, 1))
    ref = L.Standardizer()
    ref.update_mean_std(obs)
    ref.update_mean_std(obs)
    pm, ps = L.normalization_params(obs)
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == n_total
    for r in res:
        np.testing.assert_allclose(r[3], ref.mean, rtol=1e-12)
        np.testing.assert_allclose(r[4], ref.std, rtol=1e-12)
        np.testing.assert_allclose(r[5], pm, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r[6], ps, rtol=1e-10)
        np.testing.assert_allclose(r[7], [adv.mean(), adv.std(ddof=1) + 1e-5], rtol=1e-10)
        np.testing.assert_allclose(r[8], [adv.mean(), adv.std() + 1e-8], rtol=1e-10)
    single = np.stack(philox.draw(1234, np.arange(n_total, dtype=np.uint32), np.uint32(0)), axis=1)
    np.testing.assert_array_equal(np.concatenate([r[9] for r in res]), single)


def test_env_shard_partitions():
    from olympics_mujoco_b200 import distributed as D
    for n, w in ((1 << 20, 8), (4096, 4), (10, 3), (2, 4)):
        shards = [D.env_shard(n, r, w) for r in range(w)]
        assert shards[0][0] == 0 and sum(s[1] for s in shards) == n
        for a, b in zip(shards, shards[1:]):
            assert a[0] + a[1] == b[0]
        assert max(s[1] for s in shards) - min(s[1] for s in shards) <= 1
