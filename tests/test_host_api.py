"""CPU tests of the host side: the C ABI surface (library loads and exports every symbol the header
declares; no compute calls), constants derived by gca_params_init vs the oracle, the reference's
Operator / CAEnv / GridSpace contracts, host-side generators, env sharding + gloo all-gather."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gym_cellular_automata_b200 import _lib
    _lib.build_library()
    return _lib.load()


def test_abi_exports_every_declared_symbol(lib):
    from gym_cellular_automata_b200 import _lib
    header = open(os.path.join(ROOT, "include", "gca.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(gca_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gca_version() == _lib.GCA_VERSION


def _header_layout(tmp_path, structs):
    """sizeof and offsetof of every field of the given {struct name: [field names]} as gcc lays out include/gca.h."""
    import subprocess
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gca.h"', 'int main(){']
    for name, fields in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    return dict((k, int(v)) for k, v in (ln.split() for ln in subprocess.check_output([str(exe)]).decode().splitlines()))


def _ctypes_layout(pairs):
    out = {}
    for cname, cls in pairs.items():
        out[cname] = ctypes.sizeof(cls)
        for f in cls._fields_:
            out[f"{cname}.{f[0]}"] = getattr(cls, f[0]).offset
    return out


def test_abi_struct_sizes_match_header(tmp_path):
    """The ctypes mirrors (production binding AND the reference-side stub of examples/ref_binding.py, the file
    INTEGRATION.md shows) must have, field by field, the layout a C compiler gives include/gca.h."""
    import importlib.util
    from gym_cellular_automata_b200 import _lib
    spec = importlib.util.spec_from_file_location("ref_binding", os.path.join(ROOT, "examples", "ref_binding.py"))
    rb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rb)
    for pairs in ({"gca_params": _lib.GcaParams, "gca_state": _lib.GcaState, "gca_step_out": _lib.GcaStepOut,
                   "gca_inject": _lib.GcaInject, "gca_episode_stats": _lib.GcaEpisodeStats},
                  {"gca_params": rb.gca_params, "gca_state": rb.gca_state, "gca_step_out": rb.gca_step_out,
                   "gca_inject": rb.gca_inject}):
        want = _header_layout(tmp_path, {n: [f[0] for f in c._fields_] for n, c in pairs.items()})
        assert _ctypes_layout(pairs) == want
    # no field of the header may be missing from a mirror: the last field of each struct ends at sizeof
    hdr = open(os.path.join(ROOT, "include", "gca.h")).read()
    assert f"#define GCA_VERSION {_lib.GCA_VERSION}" in hdr and rb.GCA_VERSION == _lib.GCA_VERSION
    for cls in (_lib.GcaState, rb.gca_state, _lib.GcaStepOut, rb.gca_step_out):
        last = cls._fields_[-1]
        assert getattr(cls, last[0]).offset + ctypes.sizeof(last[1]) == ctypes.sizeof(cls)


def test_integration_md_shows_the_tested_binding():
    """INTEGRATION.md includes examples/ref_binding.py verbatim (tools/sync_integration.py rewrites the block)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    src = open(os.path.join(ROOT, "examples", "ref_binding.py")).read()
    assert src.strip() in doc


def test_abi_argument_errors_do_not_throw(lib):
    from gym_cellular_automata_b200 import _lib
    p = _lib.GcaParams()
    rc = lib.gca_params_init(ctypes.byref(p), 64, 64, 99, 0.12, 0.03, 0.001, -1.0, -1.0, 0.0, 0.06, 0, None)
    assert rc == -1 and b"K out of range" in lib.gca_last_error()
    rc = lib.gca_params_init(ctypes.byref(p), 4, 4, 1, 0.12, 0.03, 0.001, -1.0, -1.0, 0.0, 0.06, 0, None)
    assert rc == -2
    rc = lib.gca_env_step(None, None, None, None, None, None, None, 0, None)
    assert rc == -1
    with pytest.raises(_lib.GcaError):
        _lib.check(rc, "gca_env_step")


@pytest.mark.parametrize("size", [32, 64, 100, 256, 4096])
def test_params_init_matches_oracle_constants(lib, size):
    from oracle import alexandridis as ax, init_state as oinit, prng
    from gym_cellular_automata_b200.packed import make_params
    p = make_params(size, size, 4, 0.48, 0.12)
    c = ax.CAConstants(size)
    E = ax.EnvConstants(size, size, 0.48, 0.12)
    assert p.R == c.radius
    assert np.array_equal(np.array(list(p.ring_w)[1:c.radius + 1], np.float32), c.ring_weights)
    assert np.float32(p.dous_border) == c.dousing_weights[0, 0] and np.float32(p.dous_inner) == c.dousing_weights[1, 1]
    assert (p.age_lo, p.age_span, p.age_mult) == prng.randint_params(c.fire_age_min, c.fire_age_max)
    assert np.float32(p.t_move[4]) == E.movement_timings[4] and np.float32(p.t_shoot[1]) == E.shooting_timings[1]
    assert np.array_equal(np.array(list(p.onep_veg)[:6], np.float32), (np.float32(1) + ax.VEG_PROBS).astype(np.float32))
    assert np.array_equal(np.array(list(p.winds), np.float32), oinit.get_winds().astype(np.float32)[:, 0].reshape(-1))


def test_product_never_imports_oracle():
    """The product path must not route through the checker."""
    bad = []
    pkg = os.path.join(ROOT, "gym_cellular_automata_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "gca_oracle" in src:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_env_needs_gpu_and_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_cellular_automata_b200 import _lib
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    with pytest.raises(_lib.GcaError):
        AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=2)


# ---- reference contracts: Operator, CAEnv, GridSpace (reference tests/test_operator.py, test_ca_env.py,
#      test_grid_space.py) --------------------------------------------------------------------------------
def _identity_operator():
    from gym_cellular_automata_b200 import Operator

    class Identity(Operator):
        grid_dependant = True
        action_dependant = True
        context_dependant = True
        deterministic = True

        def __init__(self, *a, **k):
            super().__init__(*a, **k)

        def update(self, grid, action, context):
            return grid, context

    return Identity


def test_operator_contract():
    from gym_cellular_automata_b200 import GridSpace, spaces
    Identity = _identity_operator()
    gs = GridSpace(n=3, shape=(4, 4))
    op = Identity(grid_space=gs, action_space=spaces.Discrete(3), context_space=spaces.Discrete(3))
    assert isinstance(op.suboperators, tuple) and op.deterministic is True
    g = gs.sample()
    a, c = op.action_space.sample(), op.context_space.sample()
    ng, nc = op(g, a, c)
    assert gs.contains(ng) and op.context_space.contains(nc)
    assert op.seed(3) == [3]


def test_caenv_contract():
    from gym_cellular_automata_b200 import CAEnv, GridSpace, spaces
    Identity = _identity_operator()

    class MockCAEnv(CAEnv):
        def __init__(self, nrows=4, ncols=4):
            super().__init__(nrows, ncols)
            self.grid_space = GridSpace(n=3, shape=(nrows, ncols))
            self.action_space = spaces.Discrete(3)
            self.context_space = spaces.Discrete(3)
            self._MDP = Identity(self.grid_space, self.action_space, self.context_space)

        @property
        def MDP(self):
            return self._MDP

        @property
        def initial_state(self):
            return self.grid_space.sample(), self.context_space.sample()

        def _award(self):
            return 1.0

        def _is_done(self):
            self.done = self.steps_elapsed >= 2

        def _report(self):
            return {}

    env = MockCAEnv()
    obs, info = env.reset()
    assert env.grid_space.contains(obs[0])
    for _ in range(3):
        obs, reward, terminated, truncated, info = env.step(env.action_space.sample())
    assert terminated and reward == 1.0 and env.status()["steps_elapsed"] == 3
    with pytest.warns(UserWarning):
        out = env.step(0)  # graceful step after done (reference ca_env.py:50-62)
    assert out[1] == 0.0 and out[2] is True
    assert sum(env.count_cells().values()) == 16


def test_grid_space():
    from gym_cellular_automata_b200 import GridSpace
    a = GridSpace(n=3, shape=(5, 5), seed=7)
    b = GridSpace(values=[0, 1, 2], shape=(5, 5), seed=7)
    assert a == b and a.contains(a.sample()) and not a.contains(np.full((5, 5), 9))
    assert np.array_equal(GridSpace(n=3, shape=(5, 5), seed=7).sample(), GridSpace(n=3, shape=(5, 5), seed=7).sample())
    p = GridSpace(values=[0, 1, 2], probs=[0.1, 0.9, 0.0], shape=(2000,), seed=1).sample()
    assert (p == 2).sum() == 0 and 0.85 < (p == 1).mean() < 0.95
    with pytest.raises(AssertionError):
        GridSpace(n=3, shape=())


def test_spaces_shim():
    from gym_cellular_automata_b200 import spaces
    md = spaces.MultiDiscrete(np.array([[9, 2, 3]] * 4))
    s = md.sample()
    assert s.shape == (4, 3) and md.contains(s) and (s[:, 0] < 9).all()
    t = spaces.Tuple((spaces.Discrete(2), spaces.Dict({"a": spaces.Box(0, 1, shape=(3,), dtype=np.float64)})))
    x = t.sample()
    assert t.contains(x) and len(t) == 2 and list(iter(t))[0].n == 2
    b = spaces.Box(0, float("inf"), shape=(2, 2), dtype=np.float64)
    assert b.contains(b.sample())


def test_host_generators_match_literal_restatement():
    from oracle import init_state as oi
    from gym_cellular_automata_b200.forest_fire.bulldozer.utils import init_utils as pi
    assert np.array_equal(oi.get_winds(), pi.get_winds())
    for seed in (0, 5):
        a = oi.init_density(np.random.RandomState(seed), 24, 32, 2)
        assert np.array_equal(a, pi.init_density(24, 32, 2, np.random.RandomState(seed)))
        alt = oi.init_altitude(np.random.RandomState(seed), 24, 32, 2)
        assert np.allclose(alt, pi.init_altitude(24, 32, 2, np.random.RandomState(seed)), atol=1e-12)
        assert np.allclose(oi.get_slope(alt), pi.get_slope(alt), atol=1e-12)
    assert pi.create_up_to_k_mappings(2, 1)[0].tolist() == [[0, 0], [1, 0], [0, 1]]
    s = pi.get_slope(np.random.default_rng(0).random((1, 8, 8)))
    assert (s[:, 0] == 0).all() and (s[:, :, -1] == 0).all() and (s[..., 1, 1] == 0).all()


# ---- multi-GPU host logic on CPU: world_size 2, gloo --------------------------------------------------------
def _gloo_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from gym_cellular_automata_b200.distributed import gather_episode_stats, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        total = 10
        lo, hi = shard_range(total, rank, world)
        n = hi - lo
        # the per-env statistics of the rollout loop (on a GPU: rollout_stats.EpisodeStatistics, one kernel per step)
        st = {"episode_returns": torch.zeros(n), "episode_lengths": torch.zeros(n, dtype=torch.int32),
              "returned_episode_returns": torch.zeros(n), "returned_episode_lengths": torch.zeros(n, dtype=torch.int32)}
        # env e earns reward -(e+1) per step and terminates on step 3 if e is even
        for step in range(4):
            reward = -torch.arange(lo + 1, hi + 1, dtype=torch.float32)
            term = torch.tensor([(e % 2 == 0) and step == 2 for e in range(lo, hi)])
            ret_, len_ = st["episode_returns"] + reward, st["episode_lengths"] + 1
            st["returned_episode_returns"] = torch.where(term, ret_, st["returned_episode_returns"])
            st["returned_episode_lengths"] = torch.where(term, len_, st["returned_episode_lengths"])
            st["episode_returns"] = torch.where(term, torch.zeros_like(ret_), ret_)
            st["episode_lengths"] = torch.where(term, torch.zeros_like(len_), len_)
        st["keys"] = torch.arange(2 * lo, 2 * hi, dtype=torch.int64).reshape(n, 2)   # a leaf that is not 4-byte / 1-D
        g = gather_episode_stats(st)
        assert g["keys"].tolist() == [[2 * e, 2 * e + 1] for e in range(total)]
        del g["keys"]
        ret[rank] = {k: v.tolist() for k, v in g.items()}
    finally:
        dist.destroy_process_group()


def test_env_sharding_and_stats_allgather_gloo():
    import socket
    import torch.multiprocessing as mp
    from gym_cellular_automata_b200.distributed import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard_range(65536, r, 8) for r in range(8)][-1] == (57344, 65536)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == ret[1]
    g = ret[0]
    assert len(g["episode_returns"]) == 10  # ordered by global env index
    for e in range(10):
        if e % 2 == 0:
            assert g["returned_episode_returns"][e] == -3.0 * (e + 1) and g["returned_episode_lengths"][e] == 3
            assert g["episode_returns"][e] == -1.0 * (e + 1) and g["episode_lengths"][e] == 1
        else:
            assert g["episode_returns"][e] == -4.0 * (e + 1) and g["returned_episode_lengths"][e] == 0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout
    with the keys of the bench contract, honouring --steps / --warmup, and no GPU needed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--cpu-envs", "4"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["metric"] == "cell_updates_per_s" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
