"""-m gpu: the CUDA hot path (through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fmt(reports):
    return "\n".join(f"step {s}: " + " | ".join(b) for s, b in reports[:5])


@pytest.mark.parametrize("mode", ["legacy", "partitionable"])
def test_threefry_hooks(cuda_device, mode):
    from oracle import prng
    from gym_cellular_automata_b200._lib import check, current_stream, load, ptr
    m = 0 if mode == "legacy" else 1
    key = np.array([0x13198A2E, 0x03707344], dtype=np.uint32)
    kd = torch.as_tensor(key).cuda()
    for n in (1, 2, 7, 36864, 4096):
        out = torch.empty(n, dtype=torch.uint32, device="cuda")
        check(load().gca_threefry_bits(ptr(kd), n, m, ptr(out), current_stream()))
        assert np.array_equal(out.cpu().numpy(), prng.random_bits(key, n, m)), (mode, n)
    for num in (2, 5, 64):
        out = torch.empty((num, 2), dtype=torch.uint32, device="cuda")
        check(load().gca_threefry_split(ptr(kd), num, m, ptr(out), current_stream()))
        assert np.array_equal(out.cpu().numpy(), prng.split(key, num, m)), (mode, num)


def test_pack_unpack_roundtrip(cuda_device):
    from parity_util import make_pair, read_cuda_state
    env, co, E, state, info = make_pair(N=4, K=1)
    got = read_cuda_state(env)
    ctx = state["per_env_context"]
    for k in ("true_grid", "fire_age", "dousing_count", "wind_index", "key"):
        assert np.array_equal(np.asarray(ctx[k]), got[k]), k


@pytest.mark.parametrize("mode,K,hidden", [("legacy", 1, True), ("partitionable", 1, True), ("legacy", 4, True),
                                           ("partitionable", 4, False), ("legacy", 2, False)])
def test_env_step_parity(cuda_device, mode, K, hidden):
    """Stream parity: in-kernel threefry, same keys/actions/hidden layers -> identical state after
    every step (grid, fire_age, dousing, wind, key chain, position, clock, reward, done)."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=16, K=K, mode=mode, use_hidden=hidden, seed=3)
    nbad, reports, stats = lockstep(env, co, state, 240 // K + 40, np.random.default_rng(5))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 0 and stats[2] > 0, "no draws / ignitions happened: the test exercised nothing"


def test_env_step_parity_long_episode(cuda_device):
    """Whole episodes incl. burn-outs (ages run out after >= 144 CA steps) and termination."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=8, K=4, mode="legacy", use_hidden=True, seed=11)
    nbad, reports, stats = lockstep(env, co, state, 450, np.random.default_rng(2))
    assert nbad == 0, _fmt(reports)
    assert stats[3] > 0, "no burn-outs happened"


def test_rule_parity_injected_uniforms(cuda_device):
    """Rule parity: identical injected u_burn/u_grow/age_new/u_wind/wind_step on both sides."""
    from parity_util import make_pair, lockstep
    N, K = 6, 2
    env, co, E, state, info = make_pair(N=N, K=K, mode="legacy", use_hidden=True, seed=7, p_tree=0.002)
    rng = np.random.default_rng(9)

    def inject(step):
        return {"u_burn": (rng.integers(0, 1 << 23, (K, N, 64, 64, 9)) * 2.0 ** -23).astype(np.float32) * 0.35,
                "u_grow": (rng.integers(0, 1 << 23, (K, N, 64, 64)) * 2.0 ** -23).astype(np.float32),
                "age_new": rng.integers(20, 40, (K, N, 64, 64)).astype(np.int32),
                "u_wind": rng.random((K, N)).astype(np.float32),
                "wind_step": rng.integers(1, 8, (K, N)).astype(np.int32)}

    nbad, reports, stats = lockstep(env, co, state, 60, np.random.default_rng(1), inject_fn=inject)
    assert nbad == 0, _fmt(reports)
    assert stats[2] > 0 and stats[3] > 0


def test_regrowth_p_tree(cuda_device):
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=4, K=2, mode="legacy", use_hidden=False, seed=13, p_tree=0.01)
    nbad, reports, stats = lockstep(env, co, state, 80, np.random.default_rng(3))
    assert nbad == 0, _fmt(reports)


def test_dousing_heavy(cuda_device):
    """Bulldozer shoots every step: the 5x5 dousing term is active along the fire front."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=8, K=1, mode="legacy", use_hidden=True, seed=21)
    # pre-douse a band across the grid so the front must cross it
    ctx = state["per_env_context"]
    ctx["dousing_count"][:, 40:46, :] = 1
    ctx["dousing_count"][:, :, 20:24] = 1
    from parity_util import sync
    sync(env, state, as_snapshot=True)
    nbad, reports, stats = lockstep(env, co, state, 300, np.random.default_rng(4), shoot_p=1.0)
    assert nbad == 0, _fmt(reports)


@pytest.mark.parametrize("nrows,ncols,K,tma,N", [(32, 32, 2, True, 4), (128, 128, 1, True, 2), (256, 256, 1, True, 2),
                                                 (256, 256, 2, False, 1), (72, 40, 1, True, 3), (96, 80, 3, True, 2)])
def test_tiled_parity(cuda_device, nrows, ncols, K, tma, N):
    """Grids other than 64x64 go through the tiled kernel (TMA staging when W % 16 == 0, plain
    loads otherwise / on request): same bit-exact contract against the oracle."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=N, size=nrows, ncols=ncols, K=K, mode="legacy", use_hidden=True, seed=5,
                                        hidden="random", scatter_fire=0.01, use_tma=tma, fast_slope=True)
    nbad, reports, stats = lockstep(env, co, state, 40, np.random.default_rng(6))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 0 and stats[2] > 0 and stats[3] > 0


def test_tiled_partitionable_hidden_off(cuda_device):
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=2, size=128, K=2, mode="partitionable", use_hidden=False, seed=8,
                                        scatter_fire=0.01, p_tree=0.001)
    nbad, reports, stats = lockstep(env, co, state, 30, np.random.default_rng(7))
    assert nbad == 0, _fmt(reports)


def test_large_single_grid_4096(cuda_device):
    """BASELINE config 4: one 4096x4096 grid (R = 10, 21x21 heat window), two env steps."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=1, size=4096, K=1, mode="legacy", use_hidden=True, seed=2, hidden="random",
                                        scatter_fire=0.002, fast_slope=True)
    nbad, reports, stats = lockstep(env, co, state, 2, np.random.default_rng(1))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 1000
