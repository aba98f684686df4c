import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"

# Tolerance of the north star: "positions, observations and rewards must match within 1e-5 relative in
# fp32".  Relative to max(|reference|, 1) -- i.e. rtol 1e-5 with an absolute floor of 1e-5 at unit scale
# (metres, radians, m/s), because many reference values cross zero.
RTOL = 1e-5
ATOL = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def assert_close(actual, expected, name="", rtol=RTOL, atol=ATOL):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, f"{name}: shape {actual.shape} vs {expected.shape}"
    err = np.abs(actual - expected)
    tol = atol + rtol * np.abs(expected)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{name}: {bad.sum()} of {bad.size} outside tolerance; worst at {i}: "
                             f"{actual[i]!r} vs {expected[i]!r} (err {err[i]:.3e}, tol {tol[i]:.3e})")


_KNOB_DEFAULTS = dict(play_chunk=0, h1_split=-1, a3_split=-1, serial_scan=0, disc_vail2=-1, disc_pg2=-1)


@pytest.fixture
def om_knob():
    """Set a tuning / test knob of libom_b200 (``om_debug_set``: forces a kernel variant) for one test; every knob is
    back on automatic afterwards.  The library reads its OM_* environment variables only once, at load."""
    from olympics_mujoco_b200 import _lib
    lib = _lib.load()

    def set_knob(name, value):
        _lib.check(lib.om_debug_set(name.encode(), int(value)))
    yield set_knob
    for k, v in _KNOB_DEFAULTS.items():
        lib.om_debug_set(k.encode(), v)


@pytest.fixture(scope="session")
def h1_model():
    from olympics_mujoco_b200 import mjcf
    return mjcf.load_builtin("unitree_h1")


@pytest.fixture(scope="session")
def a3_model():
    from olympics_mujoco_b200 import mjcf
    return mjcf.load_builtin("stick_figure_a3")


@pytest.fixture(scope="session")
def h1_states(h1_model):
    """Dataset-shaped (qpos, qvel) samples in MJCF order, fp32-representable float64."""
    from olympics_mujoco_b200 import synthetic
    from oracle import h1 as OH
    d = synthetic.h1_walk_dataset(n_traj=2, t_raw=600, seed=3, model=h1_model)
    keys = OH.keys(h1_model)
    sample = np.stack([d[k] for k in keys], axis=1)[::3]
    qpos, qvel = OH.set_sim_state(h1_model, sample)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    return f(qpos), f(qvel)


def a3_random_states(model, n, seed=0):
    rng = np.random.default_rng(seed)
    q = np.tile(model.qpos0, (n, 1)) + rng.normal(0, 0.3, (n, model.nq))
    q[:, 0:2] = rng.uniform(-3, 3, (n, 2))
    q[:, 2] = rng.uniform(1.0, 1.5, n)
    quat = rng.normal(0, 1, (n, 4))
    q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True) * rng.uniform(0.95, 1.05, (n, 1))
    v = rng.normal(0, 2, (n, model.nv))
    f = lambda a: a.astype(np.float32).astype(np.float64)
    return f(q), f(v)
