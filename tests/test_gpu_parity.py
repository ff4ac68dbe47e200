"""-m gpu: the CUDA hot path (through the C ABI) against the CPU oracle, bit for bit."""
import os

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
                                           ("partitionable", 4, False), ("legacy", 2, False), ("legacy", 3, True),
                                           ("partitionable", 8, True)])
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
    """Grids other than 64x64 through the generic tiled kernels (TMA staging when W % 16 == 0, plain
    loads otherwise / on request): same bit-exact contract against the oracle."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=N, size=nrows, ncols=ncols, K=K, mode="legacy", use_hidden=True, seed=5,
                                        hidden="random", scatter_fire=0.01, use_tma=tma, fast_slope=True, generic_tiles=True)
    nbad, reports, stats = lockstep(env, co, state, 40, np.random.default_rng(6))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 0 and stats[2] > 0 and stats[3] > 0


@pytest.mark.parametrize("nrows,ncols,K,N,mode,hidden,p_tree", [
    (128, 128, 1, 3, "legacy", True, 0.0), (256, 256, 2, 2, "legacy", True, 0.0), (192, 192, 4, 2, "partitionable", True, 0.0),
    (256, 256, 4, 2, "legacy", False, 0.0), (64, 128, 3, 3, "legacy", True, 0.0), (128, 256, 2, 2, "partitionable", False, 0.001),
    (256, 64, 8, 2, "legacy", True, 0.0)])
def test_bitboard_grid_kernel_parity(cuda_device, nrows, ncols, K, N, mode, hidden, p_tree):
    """Grids of whole 64-bit words up to 256x256 (BASELINE config 3's grid) step with ONE launch per env step: the grid
    as tree / fire / doused bit-boards in one CTA's shared memory, all K sub-steps on-chip (csrc/gca_bb.cu).  Lock step
    with the oracle: both stream layouts, hidden layers on / off, regrowth, dousing, non-square grids."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=N, size=nrows, ncols=ncols, K=K, mode=mode, use_hidden=hidden, seed=5,
                                        hidden="random", scatter_fire=0.01, fast_slope=True, p_tree=p_tree)
    nbad, reports, stats = lockstep(env, co, state, 40 // K + 12, np.random.default_rng(6), shoot_p=0.7)
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 0 and stats[2] > 0 and stats[3] > 0


@pytest.mark.parametrize("generic", [False, True])
def test_rule_parity_injected_uniforms_128(cuda_device, generic):
    """Rule parity on a 128x128 grid with identical injected random fields on both sides: the whole-grid bit-board
    kernel and the generic tiled kernels, with regrowth, dousing bands and short injected fire ages (burn-outs inside
    the run, exercising the burn list)."""
    from parity_util import make_pair, lockstep, sync
    N, K, S = 3, 2, 128
    env, co, E, state, info = make_pair(N=N, size=S, K=K, mode="legacy", use_hidden=True, seed=7, p_tree=0.002,
                                        hidden="random", scatter_fire=0.004, fast_slope=True, generic_tiles=generic)
    state["per_env_context"]["dousing_count"][:, 60:64, :] = 1
    sync(env, state, as_snapshot=True)
    rng = np.random.default_rng(9)

    def inject(step):
        return {"u_burn": (rng.integers(0, 1 << 23, (K, N, S, S, 9)) * 2.0 ** -23).astype(np.float32) * 0.35,
                "u_grow": (rng.integers(0, 1 << 23, (K, N, S, S)) * 2.0 ** -23).astype(np.float32),
                "age_new": rng.integers(3, 12, (K, N, S, S)).astype(np.int32),
                "u_wind": rng.random((K, N)).astype(np.float32),
                "wind_step": rng.integers(1, 8, (K, N)).astype(np.int32)}

    nbad, reports, stats = lockstep(env, co, state, 24, np.random.default_rng(1), inject_fn=inject)
    assert nbad == 0, _fmt(reports)
    assert stats[2] > 0 and stats[3] > 0


def test_bitboard_grid_burn_list_overflow(cuda_device):
    """More than 1024 cells of one env burn out inside one env step (every burning cell is given a remaining age of
    1..K): the bit-board kernel drops its burn list and scans per sub-step -- still bit-exact."""
    from parity_util import make_pair, lockstep, sync
    env, co, E, state, info = make_pair(N=2, size=128, K=4, mode="legacy", use_hidden=True, seed=3, hidden="random",
                                        scatter_fire=0.12, fast_slope=True)
    ctx = state["per_env_context"]
    rs = np.random.default_rng(0)
    m = ctx["true_grid"] == 2
    assert m[0].sum() > 1200
    ctx["fire_age"][m] = rs.integers(1, 5, size=int(m.sum())).astype(np.float32)
    sync(env, state, as_snapshot=True)
    nbad, reports, stats = lockstep(env, co, state, 6, np.random.default_rng(2))
    assert nbad == 0, _fmt(reports)
    assert stats[3] > 2400


def test_tiled_partitionable_hidden_off(cuda_device):
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=2, size=128, K=2, mode="partitionable", use_hidden=False, seed=8,
                                        scatter_fire=0.01, p_tree=0.001)
    nbad, reports, stats = lockstep(env, co, state, 30, np.random.default_rng(7))
    assert nbad == 0, _fmt(reports)


def test_full_batch_4096_envs_lockstep(cuda_device):
    """BASELINE config 2 at its full size: 4096 envs of 64x64, K = 4, hidden layers on -- one full wave of 293 CTAs,
    the load balancer re-dealing the envs to CTA slots every 8 steps -- in lock step with the C oracle (which steps
    the whole batch in well under a second on the host cores): every state component of every env after every env
    step.  Scattered fires give each env its own front; size-independent properties are checked on the whole batch
    as well (counts = populations of the grid, reward = -f/(t+f+1e-8), done = no fire, only legal transitions)."""
    from parity_util import make_pair, lockstep, read_cuda_state
    N = 4096
    env, co, E, state, info = make_pair(N=N, K=4, mode="legacy", use_hidden=True, seed=9, hidden="random",
                                        scatter_fire=0.004, fast_slope=True)
    env.balance_every = 8  # as in bench.py
    before = read_cuda_state(env)["true_grid"].copy()
    nbad, reports, stats = lockstep(env, co, state, 18, np.random.default_rng(3))
    assert nbad == 0, _fmt(reports)
    assert env._state.order is not None and stats[1] > 50 * N, "no balancing / too few draws for a full-size run"
    g = read_cuda_state(env)["true_grid"]
    t, f = (g == 1).sum(axis=(1, 2)), (g == 2).sum(axis=(1, 2))
    counts = env._out.counts.cpu().numpy()
    assert np.array_equal(counts[:, 0], t) and np.array_equal(counts[:, 1], f)
    rew = -(f.astype(np.float32) / ((t + f).astype(np.float32) + np.float32(1e-8)))
    assert np.array_equal(env._out.step_reward.cpu().numpy(), rew.astype(np.float32))
    assert np.array_equal(env._out.terminated.cpu().numpy().astype(bool), f == 0)
    # p_tree = 0: empty stays empty, a tree stays or ignites (or has burnt out since), fire never reverts to tree
    assert not ((before == 0) & (g != 0)).any() and not ((before == 2) & (g == 1)).any()
    assert len({g[i].tobytes() for i in range(0, N, 16)}) == N // 16, "every env should have its own grid"


def test_two_waves_8192_envs_lockstep(cuda_device):
    """BASELINE config 5's per-GPU share (65536 envs over 8 GPUs = 8192 per GPU): 586 CTAs = two waves of the 64x64
    kernel, two chunks of the balancer.  Ten env steps in lock step with the C oracle, every env."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=8192, K=4, mode="legacy", use_hidden=True, seed=10, hidden="random",
                                        scatter_fire=0.004, fast_slope=True)
    env.balance_every = 4
    nbad, reports, stats = lockstep(env, co, state, 10, np.random.default_rng(4))
    assert nbad == 0, _fmt(reports)
    order = env._state.order.cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(8192)), "the dealing must be a permutation of the envs"


@pytest.mark.parametrize("hidden", [True, False])
def test_config3_full_batch_256(cuda_device, hidden):
    """BASELINE config 3 at its full size: 1024 envs of 256x256 (R = 6), hidden layers on and off, ten env steps of
    K = 2 in lock step with the C oracle -- every state component of every env after every step (whole-grid
    bit-board kernel, one launch per env step)."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=1024, size=256, K=2, mode="legacy", use_hidden=hidden, seed=11, hidden="random",
                                        scatter_fire=0.003, fast_slope=True)
    nbad, reports, stats = lockstep(env, co, state, 10, np.random.default_rng(5))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 1000000 and stats[2] > 1000


def test_tiled_many_envs_256(cuda_device):
    """The generic tiled kernels on config 3's grid with enough envs for several waves of tile CTAs (96 envs x 32
    tiles): three env steps of K = 2 in lock step with the C oracle."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=96, size=256, K=2, mode="legacy", use_hidden=True, seed=11, hidden="random",
                                        scatter_fire=0.003, fast_slope=True, generic_tiles=True)
    nbad, reports, stats = lockstep(env, co, state, 3, np.random.default_rng(5))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 10000


def test_large_single_grid_4096(cuda_device):
    """BASELINE config 4: one 4096x4096 grid (R = 10, 21x21 heat window), eight env steps of K = 4 CA sub-steps."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=1, size=4096, K=4, mode="legacy", use_hidden=True, seed=2, hidden="random",
                                        scatter_fire=0.002, fast_slope=True)
    nbad, reports, stats = lockstep(env, co, state, 8, np.random.default_rng(1))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 1000


# ---------------------------------------------------------------------------------------------
# observation, reset, API surface
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("enable_ext", [False, True])
def test_observation_parity(cuda_device, enable_ext):
    """RGB observation of stateless_step == oracle build_observation (new grid / position, old
    dousing marks and day/night, extension-channel quirk), float32, bit for bit."""
    from oracle import alexandridis as ax
    from parity_util import make_pair, random_actions
    N = 6
    env, co, E, state, info = make_pair(N=N, K=1, mode="legacy", use_hidden=True, seed=4, obs_mode="rgb_f32",
                                        enable_extensions=enable_ext)
    rng = np.random.default_rng(12)
    # make the day/night flip and some dousing visible
    state["per_env_context"]["time_step"][:] = np.array([398, 399, 400, 1, 799, 5], dtype=np.int32)
    state["per_env_context"]["is_night"][:] = np.array([0, 0, 1, 1, 1, 0], dtype=np.int32)
    state["per_env_context"]["dousing_count"][:, 5:9, 50:60] = 1
    state["per_env_context"]["true_grid"][3, 0, :] = 0.0  # first row empty: exercises the channel-index quirk
    from parity_util import sync
    sync(env, state, as_snapshot=True)
    o_info = {k: np.zeros(N, np.float32) for k in ("steps_elapsed", "reward_accumulated")}
    for step in range(12):
        act = random_actions(rng, N)
        rgb_o, state, reward, term, trunc, o_info = ax.stateless_step(E, act, state, o_info, K=1,
                                                                      enable_extensions=enable_ext)
        obs, r, t, tr, inf = env.stateless_step(act)
        rgb = obs[0].cpu().numpy()
        assert rgb.dtype == np.float32 and rgb.shape == (N, 64, 64, 3)
        assert np.array_equal(rgb, rgb_o), f"step {step}: {np.argwhere(rgb != rgb_o)[:4]}"
        assert np.array_equal(r.cpu().numpy(), reward)
        assert np.array_equal(inf["steps_elapsed"].cpu().numpy(), o_info["steps_elapsed"])
        assert np.array_equal(inf["reward_accumulated"].cpu().numpy(), o_info["reward_accumulated"])
        ctx = obs[1]["per_env_context"]
        assert np.array_equal(ctx["true_grid"].cpu().numpy(), state["per_env_context"]["true_grid"])
        assert np.array_equal(ctx["fire_age"].cpu().numpy(), state["per_env_context"]["fire_age"])
        assert np.array_equal(ctx["is_night"].cpu().numpy(), state["per_env_context"]["is_night"])


@pytest.mark.parametrize("nrows,ncols,mode", [(40, 36, "rgb_f32"), (30, 30, "rgb_f32"), (32, 80, "rgb_u8"), (40, 36, "rgb_u8")])
def test_observation_other_widths(cuda_device, nrows, ncols, mode):
    """The three render kernels: 16 cells per thread (width a multiple of 16: 32x80, several chunks per env), 4 cells per
    thread (40x36) and one cell per thread (30x30), float32 and uint8 pixels, extensions on, against the oracle."""
    from oracle import alexandridis as ax
    from parity_util import make_pair, random_actions, sync
    N = 5
    env, co, E, state, info = make_pair(N=N, size=nrows, ncols=ncols, K=1, mode="legacy", use_hidden=True, seed=8,
                                        obs_mode=mode, enable_extensions=True, scatter_fire=0.02)
    state["per_env_context"]["is_night"][:] = np.array([0, 1, 0, 1, 1], dtype=np.int32)
    state["per_env_context"]["dousing_count"][:, 3:6, 10:25] = 1
    state["per_env_context"]["true_grid"][2, 0, :] = 0.0
    sync(env, state, as_snapshot=True)
    o_info = {k: np.zeros(N, np.float32) for k in ("steps_elapsed", "reward_accumulated")}
    rng = np.random.default_rng(2)
    for step in range(4):
        act = random_actions(rng, N)
        rgb_o, state, reward, term, trunc, o_info = ax.stateless_step(E, act, state, o_info, K=1, enable_extensions=True)
        obs, r, t, tr, inf = env.stateless_step(act)
        rgb = obs[0].cpu().numpy()
        want = rgb_o if mode == "rgb_f32" else rgb_o.astype(np.uint8)
        assert rgb.dtype == want.dtype and rgb.shape == (N, nrows, ncols, 3)
        assert np.array_equal(rgb, want), f"step {step}: {np.argwhere(rgb != want)[:4]}"


def test_conditional_reset_parity(cuda_device):
    """stateless_step + conditional_reset against the oracle across episode ends: restored grid /
    keys / position / clock, kept time_step / is_night, zeroed info counters, recomputed reward,
    cleared terminated, re-rendered observation of the reset envs."""
    import copy
    from oracle import alexandridis as ax
    from parity_util import make_pair, random_actions, sync
    N = 6
    env, co, E, state, info = make_pair(N=N, K=4, mode="legacy", use_hidden=True, seed=9, obs_mode="rgb_f32")
    # short episodes: fires about to burn out in some envs
    ctx = state["per_env_context"]
    ctx["fire_age"][ctx["true_grid"] == 2] = np.float32(3)
    ctx["true_grid"][0:3, 40:56, 8:24] = np.where(ctx["true_grid"][0:3, 40:56, 8:24] == 1, 0, ctx["true_grid"][0:3, 40:56, 8:24])
    sync(env, state, as_snapshot=True)
    initial = copy.deepcopy(state)
    o_info = {k: np.zeros(N, np.float32) for k in ("steps_elapsed", "reward_accumulated")}
    rng = np.random.default_rng(3)
    n_resets = 0
    for step in range(10):
        act = random_actions(rng, N)
        rgb_o, state, reward, term, trunc, o_info = ax.stateless_step(E, act, state, o_info, K=4)
        tup = env.stateless_step(act)
        assert np.array_equal(tup[2].cpu().numpy(), term), step
        n_resets += int(term.sum())
        rgb_o, state, reward, term2, o_info = ax.conditional_reset(E, rgb_o, state, reward, term, o_info, act, initial)
        obs, r, t, tr, inf = env.conditional_reset(tup, act)
        assert not t.any() and not term2.any()
        assert np.array_equal(r.cpu().numpy(), reward), step
        assert np.array_equal(obs[0].cpu().numpy(), rgb_o), step
        c = obs[1]["per_env_context"]
        for k in ("true_grid", "fire_age", "dousing_count", "wind_index", "key", "is_night", "time_step"):
            assert np.array_equal(c[k].cpu().numpy(), state["per_env_context"][k]), (step, k)
        assert np.array_equal(obs[1]["position"].cpu().numpy(), state["position"])
        assert np.array_equal(obs[1]["time"].cpu().numpy(), state["time"])
        assert np.array_equal(inf["steps_elapsed"].cpu().numpy(), o_info["steps_elapsed"])
    assert n_resets > 0, "no episode ended: the reset path was not exercised"


def test_fused_auto_reset_matches_two_call_path(cuda_device):
    """GCA_FLAG_AUTO_RESET inside the step kernel == stateless_step followed by conditional_reset."""
    from parity_util import make_pair, random_actions, sync, read_cuda_state
    N = 8
    envs = []
    for fused in (False, True):
        env, co, E, state, info = make_pair(N=N, K=4, mode="legacy", use_hidden=True, seed=9)
        ctx = state["per_env_context"]
        ctx["fire_age"][ctx["true_grid"] == 2] = np.float32(2)
        sync(env, state, as_snapshot=True)
        envs.append(env)
    rng = np.random.default_rng(5)
    seen_done = 0
    for step in range(12):
        act = random_actions(rng, N)
        a = torch.as_tensor(act, device="cuda")
        tup = envs[0].stateless_step(act)
        seen_done += int(tup[2].sum())
        tup = envs[0].conditional_reset(tup, act)
        out = envs[1].step_device(a, auto_reset=True)
        s0, s1 = read_cuda_state(envs[0]), read_cuda_state(envs[1])
        for k in s0:
            assert np.array_equal(s0[k], s1[k]), (step, k)
        assert np.array_equal(tup[1].cpu().numpy(), out.reward.cpu().numpy())
    assert seen_done > 0


def test_operator_level_api(cuda_device):
    """The reference's operator call signatures (batched): CA update, RepeatCA clock, Move, Modify,
    MoveModify, MDP.update."""
    from oracle import alexandridis as ax, init_state as oinit, prng
    from parity_util import make_pair, random_actions
    N = 4
    env, co, E, state, info = make_pair(N=N, K=1, mode="legacy", use_hidden=True, seed=6)
    ctx = {k: v.copy() for k, v in state["per_env_context"].items()}
    shared = state["shared_context"]
    # CA operator
    g_o, c_o = ax.ca_update(E.ca, ctx["true_grid"], ctx, shared, prng.LEGACY)
    g, c, sh = env.ca(ctx["true_grid"], None, ctx, shared)
    assert np.array_equal(g.cpu().numpy(), g_o)
    assert np.array_equal(c["fire_age"].cpu().numpy(), c_o["fire_age"])
    assert np.array_equal(c["key"].cpu().numpy(), c_o["key"])
    assert np.array_equal(c["wind_index"].cpu().numpy(), c_o["wind_index"])
    # Move / Modify
    rng = np.random.default_rng(0)
    pos = np.stack([rng.integers(0, 64, 50), rng.integers(0, 64, 50)], 1).astype(np.int32)
    pos[:8] = [[0, 0], [0, 63], [63, 0], [63, 63], [0, 5], [5, 0], [63, 7], [7, 63]]
    for a0 in range(9):
        _, new = env.move(np.zeros((50, 64, 64), np.float32), np.full(50, a0), pos)
        assert np.array_equal(new.cpu().numpy(), ax.move(pos, np.full(50, a0), 64, 64)), a0
    dc = np.zeros((N, 64, 64), np.int32)
    p4 = pos[:N]
    _, _, pe = env.modify(np.zeros((N, 64, 64), np.float32), np.array([1, 0, 1, 1]), p4, {"dousing_count": dc})
    assert np.array_equal(pe["dousing_count"].cpu().numpy(), ax.modify(dc, np.array([1, 0, 1, 1]), p4))
    # RepeatCA clock
    acts = random_actions(rng, N)
    a0 = torch.as_tensor(acts[:, 0], device="cuda")
    a1 = torch.as_tensor(acts[:, 1], device="cuda")
    t_in = np.array([0.0, 0.5, 0.95, 0.999], dtype=np.float32)
    g2, (c2, frac) = env.repeater(ctx["true_grid"], (a0, a1), ctx, shared, torch.as_tensor(t_in, device="cuda"))
    t_o = (t_in + ((E.movement_timings[acts[:, 0]] + E.shooting_timings[acts[:, 1]]) + E.t_any_f32)).astype(np.float32)
    assert np.array_equal(frac.cpu().numpy(), np.modf(t_o)[0].astype(np.float32))
    assert np.array_equal(g2.cpu().numpy(), g_o)
    # MDP.update
    a4 = ax.full_actions(acts)
    (rgb_o, grid_o), (nctx, npos, ntime) = ax.mdp_update(E, ctx["true_grid"], a4, ctx, shared, state["position"],
                                                         state["time"], K=1, render_obs=False)
    (rgb, grid, _), (pe, position, time) = env.MDP(ctx["true_grid"], a4, ctx, shared, state["position"], state["time"])
    assert np.array_equal(grid.cpu().numpy(), grid_o)
    assert np.array_equal(position.cpu().numpy(), npos) and np.array_equal(time.cpu().numpy(), ntime)
    assert np.array_equal(pe["dousing_count"].cpu().numpy(), nctx["dousing_count"])


def test_move_douse_cuda_reproduces_reference_source_golden(cuda_device):
    """gca_move_modify (through MoveModifyCUDA) against the reference's own MoveModifyJax: every action at every cell
    of the two outer rings of a 16x16 grid (tests/golden/make_reference_golden.py run_operator_edges)."""
    import ref_golden_util as R
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    fx = R.load_case("operator_edges")
    S, a = int(fx["size"]), fx["actions"]
    n = len(a)
    env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=4, seed=0, use_hidden=False, obs_mode="none")
    grid = np.zeros((n, S, S), np.float32)
    pec = {"dousing_count": np.zeros((n, S, S), np.int32)}
    _, pos, pec = env.move_modify(grid, (a[:, 0], a[:, 1]), fx["pos_in"], pec)
    assert np.array_equal(pos.cpu().numpy(), fx["pos_out"])
    want = np.zeros((n, S, S), np.int32)
    hit = fx["doused"][:, 0] >= 0
    want[np.nonzero(hit)[0], fx["doused"][hit, 0], fx["doused"][hit, 1]] = 1
    assert np.array_equal(pec["dousing_count"].cpu().numpy(), want)


def test_env_api_surface(cuda_device):
    """reset / spaces / info keys a jax_ppo-style caller touches (reference agents/jax_ppo.py:708-735,790-791)."""
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=4, seed=0, enable_extensions=True)
    obs, info = env.reset()
    rgb, context = obs
    assert rgb.shape == (4, 64, 64, 3) and rgb.dtype == torch.float32
    assert set(info) == {"TimeLimit.truncated", "terminated", "steps_elapsed", "reward_accumulated", "reward"}
    assert env.action_space.nvec[0].tolist() == [9, 2]
    assert env.total_action_space.shape[-1] == 3 and env.extension_choices == [(2, 1)]
    grid_space, context_space = env.observation_space
    assert grid_space.shape == (4, 64, 64, 3)
    sample = env.observation_space.sample()
    assert sample[0].shape == (4, 64, 64, 3) and "per_env_context" in sample[1]
    for k in env.per_env_context_keys:
        assert k in context["per_env_context"], k
    # initial fire seeds and bulldozer position (advanced_bulldozer.py:673-700)
    g = context["per_env_context"]["true_grid"].cpu().numpy()
    assert (g[:, 48, 16] == 2).all() and (g[:, 48, 15] == 2).all() and (g == 2).sum() == 8
    assert context["position"].cpu().numpy().tolist() == [[9, 54]] * 4
    a = env.total_action_space.sample()
    obs, reward, terminated, truncated, info = env.stateless_step(a, obs, info)
    assert reward.shape == (4,) and terminated.dtype == torch.bool and not truncated.any()
    assert float(info["steps_elapsed"][0]) == 1.0
    c = env.count_cells()
    assert int(c[0] + c[1] + c[2]) == 4 * 64 * 64
    with pytest.raises(RuntimeError):  # the lazily unpacked arrays of an OLD observation are refused
        context["per_env_context"]["fire_age"]


def test_cuda_reproduces_golden_fixtures(cuda_device):
    """Committed fixtures (tests/golden/make_golden.py, oracle-generated): same seeds -> same final state."""
    import importlib.util
    import os
    from oracle import alexandridis as ax, init_state as oinit
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(here, "golden", "env_step_golden.npz"))
    for name, case in mg.CASES.items():
        mode = "legacy" if case["mode"] == 0 else "partitionable"
        state, info = oinit.initial_state(64, 64, case["N"], seed=case["seed"], jax_seed=1,
                                          use_hidden=case["use_hidden"], mode=case["mode"])
        env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=case["N"], speed_move=0.48, speed_act=0.12,
                                             use_hidden=case["use_hidden"], substeps=case["K"], rng_mode=mode,
                                             seed=0, hidden="random", obs_mode="none")
        env.set_state(state["per_env_context"], state["position"], state["time"], as_snapshot=True)
        rewards = []
        for s in range(case["steps"]):
            out = env.step_device(torch.as_tensor(mg.actions_for(case, s), device="cuda"))
            rewards.append(out.step_reward.cpu().numpy().copy())
        ref = env._state.unpack_to_reference(env._params)
        assert np.array_equal(ref["true_grid"].cpu().numpy().astype(np.uint8), gold[f"{name}/grid"]), name
        assert np.array_equal(ref["fire_age"].cpu().numpy().astype(np.uint16), gold[f"{name}/fire_age"]), name
        assert np.array_equal(np.packbits(ref["dousing_count"].cpu().numpy().astype(np.uint8), axis=-1),
                              gold[f"{name}/dousing"]), name
        assert np.array_equal(env._state.key.cpu().numpy(), gold[f"{name}/key"]), name
        assert np.array_equal(env._state.position.cpu().numpy(), gold[f"{name}/position"]), name
        assert np.array_equal(env._state.time.cpu().numpy(), gold[f"{name}/time"]), name
        assert np.array_equal(np.stack(rewards), gold[f"{name}/rewards"]), name
        assert np.array_equal(env._state.reward_accumulated.cpu().numpy(), gold[f"{name}/reward_accumulated"]), name


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", ["ref64_legacy_ext", "ref32_nohidden_regrow", "ref64_partitionable"])
def test_cuda_reproduces_reference_source_golden(cuda_device, name, fused):
    """The CUDA path against vectors recorded from the reference's OWN source (run under oracle/ref_shim,
    tests/golden/make_reference_golden.py): the reference's rollout loop -- stateless_step, then conditional_reset --
    through the mirrored env API, every state component and every float32 observation pixel of every step
    (fused=False); and the hot path proper, the fused step kernel with GCA_FLAG_AUTO_RESET (fused=True), on the
    state after each step.  64x64 cases run env_step64_kernel, the 32x32 case the tiled kernels."""
    import ref_golden_util as R
    from parity_util import read_cuda_state
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    fx = R.load_case(name)
    c = R.CASES[name]
    E, start, info, snap, slope = R.oracle_states(fx)
    N, H, W = fx["start/grid"].shape
    env = AdvancedForestFireBulldozerEnv(H, W, key=1, num_envs=N, speed_move=0.48, speed_act=0.12,
                                         use_hidden=c["use_hidden"], substeps=1,
                                         rng_mode="legacy" if c["mode"] == 0 else "partitionable", seed=0,
                                         hidden="random" if c["use_hidden"] else "reference",
                                         obs_mode="none" if fused else "rgb_f32", enable_extensions=c["ext"],
                                         ca_p_tree=float(fx["shared_scalars"][1]), auto_reset=fused)
    env.set_state(snap["per_env_context"], snap["position"], snap["time"], as_snapshot=True)
    env.set_state(start["per_env_context"], start["position"], start["time"], as_snapshot=False, info=info)
    bad = []
    for s in range(fx["actions"].shape[0]):
        a = torch.as_tensor(fx["actions"][s], device=cuda_device)
        if fused:
            out = env.step_device(a)
            reward, step_reward, term = out.reward, out.step_reward, out.terminated
            steps_elapsed, reward_acc = env._state.steps_elapsed, env._state.reward_accumulated
        else:
            step_tuple = env.stateless_step(a)
            step_reward, term = step_tuple[4]["reward"], step_tuple[2]
            if R.sha(step_tuple[0][0].cpu().numpy().astype(np.float32)) != str(fx["steps/pre_rgb_sha256"][s]):
                bad.append(f"{name} step {s}: observation before conditional_reset differs")
            obs, reward, term_after, _, ninfo = env.conditional_reset(step_tuple, a)
            steps_elapsed, reward_acc = ninfo["steps_elapsed"], ninfo["reward_accumulated"]
            assert not term_after.any()
            if R.sha(obs[0].cpu().numpy().astype(np.float32)) != str(fx["steps/rgb_sha256"][s]):
                bad.append(f"{name} step {s}: observation differs")
        st = read_cuda_state(env)
        got = {"grid": st["true_grid"].astype(np.uint8), "fire_age": st["fire_age"].astype(np.uint16),
               "dousing": np.packbits(st["dousing_count"].astype(np.uint8), axis=-1), "key": st["key"],
               "wind_index": st["wind_index"].astype(np.int32), "time_step": st["time_step"].astype(np.int32),
               "is_night": st["is_night"].astype(np.int32), "position": st["position"].astype(np.int32),
               "time": st["time"].astype(np.float32), "step_reward": step_reward.cpu().numpy(),
               "terminated": term.cpu().numpy().astype(np.uint8), "reward": reward.cpu().numpy(),
               "steps_elapsed": steps_elapsed.cpu().numpy(), "reward_accumulated": reward_acc.cpu().numpy()}
        bad += R.compare_step(fx, s, got, name)
    assert not bad, "\n".join(bad[:20])


def test_env_keys_do_not_depend_on_sharding(cuda_device):
    from oracle import prng
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    kw = dict(use_hidden=False, obs_mode="none", seed=0)
    whole = AdvancedForestFireBulldozerEnv(64, 64, key=7, num_envs=8, **kw)
    whole.reset()
    parts = [AdvancedForestFireBulldozerEnv(64, 64, key=7, num_envs=4, env_offset=o, total_envs=8, **kw) for o in (0, 4)]
    for p in parts:
        p.reset()
    keys = torch.cat([p._state.key for p in parts]).cpu().numpy()
    assert np.array_equal(whole._state.key.cpu().numpy(), keys)
    assert np.array_equal(keys, prng.split(prng.key_from_seed(7), 8, prng.LEGACY))


def test_load_balancing_is_a_permutation_and_changes_nothing(cuda_device):
    """gca_balance_order only re-deals envs to warps: states stay bit-exact vs the oracle."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=300, K=2, mode="legacy", use_hidden=False, seed=4)
    env.balance_every = 3
    nbad, reports, stats = lockstep(env, co, state, 14, np.random.default_rng(8))
    assert nbad == 0, _fmt(reports)
    order = env._state.order.cpu().numpy()
    assert sorted(order.tolist()) == list(range(300))
    work = env._state.work.cpu().numpy()
    assert work.max() > 0


# ---------------------------------------------------------------------------------------------
# v3 rule set (WindyForestFire + NumPy Move / Modify / RepeatCA semantics)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W", [(16, 16), (64, 96), (256, 256), (40, 130)])
def test_windy_ca_rule_parity(cuda_device, H, W):
    from oracle import windy
    from gym_cellular_automata_b200.forest_fire.operators import WindyForestFire
    rng = np.random.default_rng(H * 7 + W)
    ca = WindyForestFire()
    N = 3
    g = rng.choice([0, 3, 25], size=(N, H, W), p=[0.15, 0.8, 0.05]).astype(np.int64)
    wind = windy.DEFAULT_WIND
    for step in range(6):
        roll = rng.random((N, 3, 3))
        exp = np.stack([windy.windy_update(g[e], wind, roll[e]) for e in range(N)])
        got, _ = ca(g, None, wind, roll=roll)
        assert np.array_equal(got.cpu().numpy(), exp), (step, np.argwhere(got.cpu().numpy() != exp)[:5])
        g = exp
    # the reference's deterministic case: wind = 1 -> every burning neighbour propagates
    got, _ = ca(g[0], None, np.ones((3, 3)), roll=rng.random((1, 3, 3)))
    assert np.array_equal(got.cpu().numpy(), windy.windy_update(g[0], np.ones((3, 3)), np.zeros((3, 3))))


def test_v3_env_step_parity(cuda_device):
    """Batched v3 env step == the oracle's per-env CAEnv.step: float64 clock with 0, 1 or several CA
    updates per step, clamped move, tree cut, reward, done."""
    from oracle import windy
    from gym_cellular_automata_b200.forest_fire.bulldozer import ForestFireBulldozerEnv
    N, H, W = 5, 48, 80
    env = ForestFireBulldozerEnv(H, W, num_envs=N, seed=3, t_move=0.45, t_shoot=0.8, max_repeats=3)
    obs, info = env.reset()
    grid = obs[0].cpu().numpy().copy()
    pos = obs[1][1].cpu().numpy().copy()
    time = obs[1][2].cpu().numpy().copy()
    assert (grid == 25).sum() == N and set(np.unique(grid)) <= {0, 3, 25}
    C = windy.V3Constants(H, W, t_move=0.45, t_shoot=0.8)
    rng = np.random.default_rng(1)
    seen = set()
    for step in range(40):
        act = np.stack([rng.integers(0, 9, N), rng.integers(0, 2, N)], 1)
        rolls = rng.random((N, 3, 9))
        obs, reward, term, trunc, info = env.step(act, rolls=rolls)
        for e in range(N):
            g, p, t, r, d, rep = windy.v3_env_step(C, grid[e], pos[e], time[e], act[e], windy.DEFAULT_WIND,
                                                   rolls[e].reshape(3, 3, 3))
            grid[e], pos[e], time[e] = g, p, t
            seen.add(rep)
            assert int(info["repeats"][e]) == rep
            rr = float(reward[e])
            assert (np.isnan(rr) and np.isnan(r)) or rr == r, (step, e, rr, r)
            assert bool(term[e]) == d
        assert np.array_equal(obs[0].cpu().numpy(), grid), step
        assert np.array_equal(obs[1][1].cpu().numpy(), pos) and np.array_equal(obs[1][2].cpu().numpy(), time)
    assert {0, 1} <= seen and max(seen) >= 2


# ---------------------------------------------------------------------------------------------
# rollout-side episode statistics (agents/jax_ppo.py:504-655)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [5, 300, 4099])
def test_episode_statistics_parity(cuda_device, N):
    """gca_episode_stats_update against the literal serial-scan restatement, over many steps with bursts
    of more than 10 simultaneous finishes (ring wrap) and truncations."""
    from gym_cellular_automata_b200.rollout_stats import EpisodeStatistics
    from oracle import rollout
    rng = np.random.default_rng(N)
    dev_st = EpisodeStatistics(N, cuda_device)
    st = rollout.new_stats(N)
    for step in range(40):
        p_fin = [0.0, 0.02, 0.5, 1.0][step % 4]
        actions = rng.integers(0, 3, (N, 3)).astype(np.int32)
        reward = (-rng.random(N)).astype(np.float32)
        term = (rng.random(N) < p_fin).astype(np.uint8)
        trunc = ((rng.random(N) < 0.01) & (term == 0)).astype(np.uint8)
        night = rng.integers(0, 2, N).astype(np.uint8)
        st = rollout.update(st, actions, reward, term, trunc, night)
        dev_st.update(torch.as_tensor(actions, device=cuda_device), torch.as_tensor(reward, device=cuda_device),
                      torch.as_tensor(term, device=cuda_device), torch.as_tensor(night, device=cuda_device),
                      torch.as_tensor(trunc, device=cuda_device))
        for k, v in dev_st.as_dict().items():
            got = v.cpu().numpy()
            want = np.asarray(st[k]).reshape(got.shape)
            assert np.array_equal(got, want), f"step {step}: {k} differs"


def test_episode_statistics_reproduce_reference_source_golden(cuda_device):
    """gca_episode_stats_update against vectors recorded from the reference's own step_env_wrapped source
    (tests/golden/make_reference_golden.py run_rollout_stats)."""
    import ref_golden_util as R
    from gym_cellular_automata_b200.rollout_stats import EpisodeStatistics
    fx = R.load_case("rollout_stats")
    steps, N = fx["actions"].shape[:2]
    dev_st = EpisodeStatistics(N, cuda_device)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=cuda_device)  # noqa: E731
    for s in range(steps):
        dev_st.update(t(fx["actions"][s]), t(fx["reward"][s]), t(fx["terminated"][s]), t(fx["is_night"][s]),
                      t(fx["truncated"][s]))
        for k, v in dev_st.as_dict().items():
            got = v.cpu().numpy()
            assert np.array_equal(got, fx["out/" + k][s].reshape(got.shape)), f"step {s}: {k} differs"


def test_episode_statistics_with_env(cuda_device):
    """Driven by the env's own outputs (step_reward / terminated / obs_night), as the rollout loop would."""
    from gym_cellular_automata_b200.rollout_stats import EpisodeStatistics
    from oracle import rollout
    from parity_util import make_pair
    env, co, E, state, info = make_pair(N=16, K=4, mode="legacy", use_hidden=True, seed=2)
    env.auto_reset = True
    dev_st = EpisodeStatistics(16, cuda_device)
    st = rollout.new_stats(16)
    rng = np.random.default_rng(3)
    for step in range(30):
        a = np.stack([rng.integers(0, 9, 16), rng.integers(0, 2, 16), rng.integers(0, 3, 16)], 1).astype(np.int32)
        ad = torch.as_tensor(a, device=cuda_device)
        night_before = env._state.is_night.cpu().numpy().copy()
        if step % 2:
            out = env.step_device(ad)
            dev_st.update(ad, out.step_reward, out.terminated, out.obs_night)
        else:  # host rollout loop: pinned actions, read in place by the step kernel and by the statistics kernel
            if step == 0:
                h_rew, h_term = env.host_result_buffers()
            ah = torch.as_tensor(a).pin_memory()
            env.step_host(ah, h_rew, h_term)
            out = env._out
            dev_st.update(ah, out.step_reward, out.terminated, out.obs_night)
            torch.cuda.synchronize()  # ah is dropped at the end of the iteration
        st = rollout.update(st, a, out.step_reward.cpu().numpy(), out.terminated.cpu().numpy(), np.zeros(16, np.uint8),
                            night_before)
    for k, v in dev_st.as_dict().items():
        assert np.array_equal(v.cpu().numpy(), np.asarray(st[k]).reshape(v.shape)), k


def test_v3_cuda_reproduces_reference_source_golden(cuda_device):
    """The batched CUDA v3 env against vectors recorded from the reference's own v3 operators
    (tests/golden/make_reference_golden.py run_v3): grid, position, float64 clock, reward, done, CA-update count."""
    import ref_golden_util as R
    from gym_cellular_automata_b200.forest_fire.bulldozer import ForestFireBulldozerEnv
    fx = R.load_case("v3_32x48")
    N, H, W = fx["grid0"].shape
    tm, ts, ta = [float(x) for x in fx["t_move_shoot_any"]]
    assert not fx["frozen"].any(), "the fixture has no finished env (a finished reference env stops stepping)"
    env = ForestFireBulldozerEnv(H, W, num_envs=N, seed=0, t_move=tm, t_shoot=ts, t_any=ta, max_repeats=3)
    env.reset()
    env.set_state(fx["grid0"].astype(np.int64), fx["position0"], fx["time0"])
    for s in range(fx["actions"].shape[0]):
        obs, reward, term, trunc, info = env.step(fx["actions"][s], rolls=fx["rolls"][s].reshape(N, 3, 9))
        assert np.array_equal(obs[0].cpu().numpy(), fx["steps/grid"][s]), s
        assert np.array_equal(obs[1][1].cpu().numpy(), fx["steps/position"][s]), s
        assert np.array_equal(obs[1][2].cpu().numpy(), fx["steps/time"][s]), s
        assert np.array_equal(reward.cpu().numpy(), fx["steps/reward"][s], equal_nan=True), s
        assert np.array_equal(term.cpu().numpy(), fx["steps/terminated"][s].astype(bool)), s
        assert np.array_equal(np.asarray(info["repeats"].cpu() if torch.is_tensor(info["repeats"]) else info["repeats"]),
                              fx["steps/repeats"][s]), s


@pytest.mark.parametrize("size", [64, 32])
@pytest.mark.parametrize("transport", ["zero_copy", "staged", "pageable"])
def test_step_host_equals_step_device(cuda_device, transport, size):
    """gca_env_step_host (host buffers in / out) gives the states and results of the device call, on each of its
    transports: pinned buffers read / written by the kernel itself (zero-copy), pinned buffers staged through the
    copy engine (GCA_FLAG_HOST_COPY), and pageable buffers (falls back to staging by itself).  Envs terminate and
    are auto-reset inside the run (reward of the restored grid goes to the host mirror as well)."""
    from parity_util import make_pair
    # size 32 runs the tiled kernels, whose host entry point always stages (pinned buffers included)
    envs = [make_pair(N=33, K=4, mode="legacy", use_hidden=True, seed=6, size=size)[0] for _ in range(2)]
    for env in envs:
        env.auto_reset = True
    rng = np.random.default_rng(1)
    pin = (lambda t: t.pin_memory()) if transport != "pageable" else (lambda t: t)
    h_rew = pin(torch.full((33,), -7.0, dtype=torch.float32))
    h_term = pin(torch.full((33,), 9, dtype=torch.uint8))
    for step in range(25):
        a = torch.as_tensor(np.stack([rng.integers(0, 9, 33), rng.integers(0, 2, 33), rng.integers(0, 3, 33)], 1).astype(np.int32))
        out = envs[0].step_device(a.to(cuda_device))
        envs[1].step_host(pin(a), h_rew, h_term, staged=(transport == "staged"))
        assert torch.equal(out.reward.cpu(), h_rew) and torch.equal(out.terminated.cpu(), h_term)
        assert torch.equal(envs[1]._out.reward.cpu(), h_rew) and torch.equal(envs[1]._out.terminated.cpu(), h_term)
    for f in ("cell", "death", "doused", "key", "position", "time", "tick", "wind_index"):
        assert torch.equal(getattr(envs[0]._state, f), getattr(envs[1]._state, f)), f


@pytest.mark.parametrize("mixed", [False, True])
def test_dense_front_multi_pass(cuda_device, mixed):
    """Hundreds of scattered fires: more than 256 front cells per env, so the 64x64 kernel rebuilds its front
    list per sub-step and works through it in several passes -- with `mixed`, next to sparse envs of the
    same CTA (the pass loop is CTA-uniform).  Short remaining ages make burn-outs start at once."""
    from parity_util import make_pair, lockstep, sync
    env, co, E, state, info = make_pair(N=20, K=4, mode="legacy", use_hidden=True, seed=17, scatter_fire=0.03)
    if mixed:
        _, _, _, plain, _ = make_pair(N=20, K=4, mode="legacy", use_hidden=True, seed=17)
        for k in ("true_grid", "fire_age"):
            state["per_env_context"][k][::2] = plain["per_env_context"][k][::2]
        sync(env, state, as_snapshot=True)
    front0 = env.stats()[0]
    nbad, reports, stats = lockstep(env, co, state, 25, np.random.default_rng(12))
    assert nbad == 0, _fmt(reports)
    assert stats[0] - front0 > 20 * 256 * (1 if mixed else 2), "the fronts were not dense enough to need several passes"
    assert stats[3] > 0


@pytest.mark.parametrize("N", [1, 13, 15, 29])
def test_env_counts_around_the_cta_size(cuda_device, N):
    """Batch sizes below, at and just past the 14 envs a CTA of the 64x64 kernel steps (warps without an env
    still join its barriers and pooled work)."""
    from parity_util import make_pair, lockstep
    env, co, E, state, info = make_pair(N=N, K=4, mode="legacy", use_hidden=True, seed=23 + N, scatter_fire=0.002)
    nbad, reports, stats = lockstep(env, co, state, 40, np.random.default_rng(N))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 0


@pytest.mark.parametrize("inject_ages", [False, True])
def test_ignition_burst_overflows_deferred_age_list(cuda_device, inject_ages):
    """Every draw ignites (injected u_burn = 0): hundreds of ignitions per env step, more than the 64x64 kernel
    buffers for its end-of-step batch of fire-age draws, so the early-flush path runs (with the in-kernel
    randint, or with injected ages)."""
    from parity_util import make_pair, lockstep
    N, K = 5, 4
    env, co, E, state, info = make_pair(N=N, K=K, mode="legacy", use_hidden=True, seed=31)
    rng = np.random.default_rng(5)

    def inject(step):
        d = {"u_burn": np.zeros((K, N, 64, 64, 9), np.float32)}
        if inject_ages:
            d["age_new"] = rng.integers(150, 160, (K, N, 64, 64)).astype(np.int32)
        return d

    ign0 = env.stats()[2]
    nbad, reports, stats = lockstep(env, co, state, 8, np.random.default_rng(1), inject_fn=inject)
    assert nbad == 0, _fmt(reports)
    assert (stats[2] - ign0) / (8 * N) > 200, "not enough ignitions per env step to overflow the deferred list"


# ---------------------------------------------------------------------------------------------
# device-side hidden-layer generation (init_utils.py:10-116,166-200)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W", [(64, 64), (32, 80), (256, 256)])
def test_generate_hidden_matches_numpy_mirror(cuda_device, H, W):
    """gca_generate_hidden against oracle/hidden_device.py: vegetation / density bit for bit, altitude to 1e-12,
    slope and slope factor to float32 rounding; and an env's layers do not depend on batch size or offset."""
    from gym_cellular_automata_b200._lib import check, current_stream, load, ptr
    from gym_cellular_automata_b200.forest_fire.bulldozer.utils.init_utils import p_slope_table
    from oracle import hidden_device as hd
    seed, N, off = 0xABCDEF0123, 5, 3

    def gen(n, offset):
        t = dict(veg=torch.empty((n, H, W), dtype=torch.int32, device=cuda_device),
                 den=torch.empty((n, H, W), dtype=torch.int32, device=cuda_device),
                 alt=torch.empty((n, H, W), dtype=torch.float32, device=cuda_device),
                 alt64=torch.empty((n, H, W), dtype=torch.float64, device=cuda_device),
                 slope=torch.empty((n, H, W, 3, 3), dtype=torch.float32, device=cuda_device),
                 pslope=torch.empty((n, H, W, 3, 3), dtype=torch.float32, device=cuda_device))
        check(load().gca_generate_hidden(n, H, W, seed, offset, ptr(t["veg"]), ptr(t["den"]), ptr(t["alt"]), ptr(t["alt64"]),
                                         ptr(t["slope"]), ptr(t["pslope"]), current_stream()), "gca_generate_hidden")
        return {k: v.cpu().numpy() for k, v in t.items()}

    g = gen(N, off)
    for e in (0, N - 1):
        assert np.array_equal(g["veg"][e], hd.patches(H, W, seed, off + e, hd.VEG_RECT, hd.VEG_FILL))
        assert np.array_equal(g["den"][e], hd.patches(H, W, seed, off + e, hd.DEN_RECT, hd.DEN_FILL))
        alt = hd.altitude(H, W, seed, off + e)
        assert np.allclose(g["alt64"][e], alt, rtol=0, atol=1e-12)
        assert np.array_equal(g["alt"][e], g["alt64"][e].astype(np.float32))
        s = hd.slope(g["alt64"][e])
        assert np.allclose(g["slope"][e], s, rtol=3e-7, atol=1e-6)
        assert np.allclose(g["pslope"][e], p_slope_table(g["slope"][e]), rtol=3e-7, atol=0)
    one = gen(1, off + 2)
    for k in ("veg", "den", "alt64", "slope", "pslope"):
        assert np.array_equal(one[k][0], g[k][2]), k


def test_env_with_device_generated_layers(cuda_device):
    """hidden="device": the env steps on layers it generated on the GPU, and the oracle -- fed the very same
    layers and slope-factor table, read back from the device -- stays bit-exact."""
    from oracle import alexandridis as ax
    from oracle import init_state as oinit
    from oracle import prng
    from oracle.c_oracle import COracle
    from parity_util import lockstep, sync
    N, K = 6, 4
    env = AdvancedForestFireBulldozerEnv_(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True,
                                          substeps=K, rng_mode="legacy", seed=5, hidden="device", obs_mode="none",
                                          collect_stats=True)
    state, info = oinit.initial_state(64, 64, N, seed=5, jax_seed=1, use_hidden=True, mode=prng.LEGACY, hidden="random")
    ctx = state["per_env_context"]
    ctx["vegetation"] = env._vegitation.cpu().numpy()
    ctx["density"] = env._density.cpu().numpy()
    ctx["altitude"] = env._altitude.cpu().numpy()
    ctx["slope"] = env._slope.cpu().numpy()
    ctx["pslope"] = env._pslope_dev.cpu().numpy()
    E = ax.EnvConstants(64, 64, speed_move=0.48, speed_act=0.12)
    winds = oinit.get_winds()
    state["shared_context"] = E.shared_context(winds)
    co = COracle(E, winds, K=K, mode=prng.LEGACY)
    sync(env, state, as_snapshot=True)
    nbad, reports, stats = lockstep(env, co, state, 40, np.random.default_rng(2))
    assert nbad == 0, _fmt(reports)
    assert stats[1] > 0


def AdvancedForestFireBulldozerEnv_(*a, **k):
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    return AdvancedForestFireBulldozerEnv(*a, **k)


def test_reference_side_stub_steps_an_env(cuda_device):
    """The binding INTEGRATION.md shows (examples/ref_binding.py: ctypes + include/gca.h only, nothing of the package
    imported) packs a reference-layout state, steps 64x64 envs with the fused kernel and unpacks them again -- in lock
    step with the oracle, auto-reset included."""
    import importlib.util
    import os
    from oracle import alexandridis as ax, init_state as oinit, prng
    from oracle.c_oracle import COracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ref_binding", os.path.join(root, "examples", "ref_binding.py"))
    rb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rb)
    lib = rb.load()
    N, K = 12, 2
    state, _ = oinit.initial_state(64, 64, N, seed=4, jax_seed=9, use_hidden=True, mode=prng.LEGACY, hidden="random")
    E = ax.EnvConstants(64, 64, speed_move=0.48, speed_act=0.12)
    winds = oinit.get_winds()
    state["shared_context"] = E.shared_context(winds)
    co = COracle(E, winds, K=K, mode=prng.LEGACY)
    env = rb.RefSideEnv(lib, torch, 64, 64, N, 0.48, 0.12, substeps=K, rng_mode=rb.GCA_RNG_LEGACY, auto_reset=False)
    ctx = state["per_env_context"]
    dev = {k: torch.as_tensor(np.ascontiguousarray(ctx[k])).cuda() for k in
           ("true_grid", "fire_age", "dousing_count", "vegetation", "density", "wind_index", "is_night", "time_step")}
    dev["vegetation"], dev["density"] = dev["vegetation"].to(torch.int32), dev["density"].to(torch.int32)
    dev["key"] = torch.as_tensor(np.asarray(ctx["key"]).astype(np.uint32)).cuda()
    ps = torch.as_tensor(np.asarray(ctx["pslope"], np.float32).reshape(N, 64, 64, 9)[..., [0, 1, 2, 3, 5, 6, 7, 8]].copy()).cuda()
    env.reset(dev, torch.as_tensor(np.asarray(state["position"], np.int32)).cuda(),
              torch.as_tensor(np.asarray(state["time"], np.float32)).cuda(), pslope8=ps)
    rng = np.random.default_rng(0)
    from parity_util import random_actions
    for step in range(80):
        act = random_actions(rng, N)
        reward, term, counts = co.step(state, act)
        r, t = env.stateless_step(torch.as_tensor(act).cuda())
        got = env.context()
        for k in ("true_grid", "fire_age", "dousing_count"):
            assert np.array_equal(got[k].cpu().numpy(), np.asarray(ctx[k])), (step, k)
        assert np.array_equal(env.step_reward.cpu().numpy(), reward), step
        assert np.array_equal(t.cpu().numpy().astype(bool), term), step
        assert np.array_equal(env.state.t["key"].cpu().numpy(), np.asarray(ctx["key"])), step
        assert np.array_equal(env.state.t["position"].cpu().numpy(), np.asarray(state["position"])), step
    rgb = env.observation(torch.as_tensor(act).cuda())
    assert rgb.shape == (N, 64, 64, 3) and float(rgb.max()) <= 255.0


def test_step_host_async_two_groups_equals_sync(cuda_device):
    """EnvPool-style loop: two env groups on two CUDA streams, step_host(wait=False) + step_host_wait(), the results of
    one group handled while the other steps -- same rewards / terminations / states as one synchronous env."""
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    N, K, steps = 64, 4, 40

    def make(n, src=None, sl=None):
        e = AdvancedForestFireBulldozerEnv(64, 64, key=3, num_envs=n, speed_move=0.48, speed_act=0.12, use_hidden=True,
                                           substeps=K, seed=5, hidden="random", obs_mode="none", auto_reset=True,
                                           balance_every=4)
        e.reset()
        if src is not None:  # this group = envs `sl` of the whole batch: same grids, layers, keys, winds
            ic = src.initial_state[1]
            e.set_state({k: np.asarray(v)[sl] for k, v in ic["per_env_context"].items()}, ic["position"][sl], ic["time"][sl],
                        as_snapshot=True)
        return e
    whole = make(N)
    rng = np.random.default_rng(1)
    acts = torch.as_tensor(np.stack([rng.integers(0, 9, (steps, N)), rng.integers(0, 2, (steps, N)),
                                     rng.integers(0, 3, (steps, N))], -1).astype(np.int32)).pin_memory()
    hr, ht = whole.host_result_buffers()
    ref_r, ref_t = [], []
    for i in range(steps):
        whole.step_host(acts[i], hr, ht)
        ref_r.append(hr.clone()); ref_t.append(ht.clone())
    halves = [make(N // 2, whole, slice(0, N // 2)), make(N // 2, whole, slice(N // 2, N))]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    bufs = [h.host_result_buffers() for h in halves]
    hacts = [acts[:, :N // 2].contiguous().pin_memory(), acts[:, N // 2:].contiguous().pin_memory()]
    got_r = [[None] * steps for _ in range(2)]
    got_t = [[None] * steps for _ in range(2)]
    for g in range(2):
        with torch.cuda.stream(streams[g]):
            halves[g].step_host(hacts[g][0], bufs[g][0], bufs[g][1], wait=False)
    for i in range(steps):
        for g in range(2):
            halves[g].step_host_wait()
            got_r[g][i], got_t[g][i] = bufs[g][0].clone(), bufs[g][1].clone()
            if i + 1 < steps:
                with torch.cuda.stream(streams[g]):
                    halves[g].step_host(hacts[g][i + 1], bufs[g][0], bufs[g][1], wait=False)
    torch.cuda.synchronize()
    for i in range(steps):
        assert torch.equal(torch.cat([got_r[0][i], got_r[1][i]]), ref_r[i]), i
        assert torch.equal(torch.cat([got_t[0][i], got_t[1][i]]), ref_t[i]), i
    assert torch.equal(torch.cat([halves[0]._state.cell, halves[1]._state.cell]), whole._state.cell)


def test_env_context_round_trips_through_the_operator_api(cuda_device):
    """The env's own (lazily unpacked) per_env_context fed back into set_state / MDP / ca / repeater -- the reference's
    operator call pattern -- keeps its lazy arrays (dict(ctx) would drop true_grid / fire_age / dousing_count)."""
    from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
    for hidden in ("reference", "device"):
        env = AdvancedForestFireBulldozerEnv(64, 64, key=2, num_envs=4, speed_move=0.48, speed_act=0.12, use_hidden=True,
                                             seed=1, hidden=hidden, obs_mode="rgb_f32")
        obs, info = env.reset()
        ctx = obs[1]
        pe = ctx["per_env_context"]
        assert set(dict(pe.items())) >= {"true_grid", "fire_age", "dousing_count", "wind_index", "key"}
        a = torch.as_tensor(np.array([[4, 1, 0]] * 4, dtype=np.int32)).cuda()
        (rgb, grid, _), (pe2, pos2, time2) = env.MDP(pe["true_grid"], a, pe, ctx["shared_context"], ctx["position"], ctx["time"])
        assert rgb.shape == (4, 64, 64, 3) and grid.shape == (4, 64, 64)
        # the returned context goes straight back in, as does the one of a plain stateless_step
        env.set_state(pe2, pos2, time2)
        obs, r, t, tr, info = env.stateless_step(a)
        env.set_state(obs[1]["per_env_context"], obs[1]["position"], obs[1]["time"])
        g2, pe3, _ = env.ca(obs[1]["per_env_context"]["true_grid"], a, obs[1]["per_env_context"], obs[1]["shared_context"])
        assert g2.shape == (4, 64, 64) and "fire_age" in pe3


@pytest.mark.parametrize("mode", ["rgb_f32", "rgb_u8"])
def test_fused_observation_and_auto_reset_frame(cuda_device, mode):
    """GCA_FLAG_RENDER (the step kernel draws the frame itself) with the fused auto-reset: every frame equals what
    stateless_step + conditional_reset of the two-call path return (stand-alone render kernel, pixel-checked against
    the oracle in test_observation_parity / test_conditional_reset_parity) -- incl. the frames of envs that reset in
    the step (restored grid / position, post-step dousing marks and day/night)."""
    from parity_util import make_pair, random_actions, sync
    from gym_cellular_automata_b200 import _lib
    N = 10
    envs = []
    for fused in (False, True):
        env, co, E, state, info = make_pair(N=N, K=2, mode="legacy", use_hidden=True, seed=13, obs_mode=mode)
        ctx = state["per_env_context"]
        ctx["fire_age"][ctx["true_grid"] == 2] = np.float32(3)   # fires die quickly: envs terminate and reset
        ctx["time_step"][:] = np.arange(395, 395 + N, dtype=np.int32)  # day/night flips inside the run
        ctx["dousing_count"][:, 8:12, 50:58] = 1
        sync(env, state, as_snapshot=True)
        envs.append(env)
    two_call, fused_env = envs
    two_call._can_fuse_render = lambda: False   # force the stand-alone render kernel on the two-call side
    fused_env.auto_reset = True
    rng = np.random.default_rng(3)
    resets = 0
    for step in range(14):
        act = random_actions(rng, N, shoot_p=0.8)
        tup = two_call.stateless_step(act)
        resets += int(tup[2].sum())
        tup = two_call.conditional_reset(tup, act)
        obs, r, t, tr, inf = fused_env.stateless_step(act)
        assert obs[0].dtype == (torch.uint8 if mode == "rgb_u8" else torch.float32)
        assert torch.equal(obs[0], tup[0][0]), (step, torch.nonzero(obs[0] != tup[0][0])[:4])
        assert torch.equal(r, tup[1])
    assert resets > 0
    # the device-level call returns the same frame as stateless_step would
    a = torch.as_tensor(random_actions(rng, N), device="cuda")
    out, rgb = fused_env.step_observe_device(a)
    tup = two_call.conditional_reset(two_call.stateless_step(a), a)
    assert torch.equal(rgb, tup[0][0])


def test_stationary_preroll_mixes_phases_and_keeps_the_state_consistent(cuda_device):
    """bench.py's workload preparation (workload.stationary_preroll: staggered forced resets through
    gca_conditional_reset between fused steps): afterwards the envs' episode ages are spread over the horizon, and the
    packed state is still self-consistent -- the bit-board twin equals the u8 grid, the row minima bound the burn-out
    ticks, and a lock-step run against the oracle from that state stays bit-exact."""
    from parity_util import make_pair, lockstep, read_cuda_state
    from gym_cellular_automata_b200.workload import stationary_preroll
    N, K = 64, 4
    env, co, E, state, info = make_pair(N=N, K=K, mode="legacy", use_hidden=True, seed=21, hidden="random", fast_slope=True)
    env.auto_reset = True
    env.balance_every = 4
    stationary_preroll(env, horizon=64, groups=8, seed=1, settle=8)
    age = env._state.steps_elapsed.cpu().numpy()
    assert len(np.unique(age)) >= 8 and age.max() <= 72 and age.min() >= 8
    st = env._state
    cell = st.cell.cpu().numpy()
    bb = st.bb.cpu().numpy().view(np.uint64).reshape(N, 64, 2)
    cols = np.arange(64, dtype=np.uint64)
    tree = ((bb[:, :, 0:1] >> cols) & np.uint64(1)).astype(bool)
    fire = ((bb[:, :, 1:2] >> cols) & np.uint64(1)).astype(bool)
    assert np.array_equal(tree, cell == 1) and np.array_equal(fire, cell == 2)
    # continue in lock step with the oracle from the state the pre-roll left
    got = read_cuda_state(env)
    ctx = state["per_env_context"]
    for k in ("true_grid", "fire_age", "dousing_count", "wind_index", "key", "is_night", "time_step"):
        ctx[k] = got[k].copy()
    state["position"], state["time"] = got["position"].copy(), got["time"].copy()
    env.auto_reset = False
    nbad, reports, stats = lockstep(env, co, state, 12, np.random.default_rng(4), resync=False)
    assert nbad == 0, _fmt(reports)


@pytest.mark.parametrize("N", [1, 13, 15, 29])
def test_fused_observation_env_counts_around_the_cta_size(cuda_device, N):
    """The fused frame for batches that do not fill the step kernel's 14-env CTAs (idle warp slots), uint8."""
    from parity_util import make_pair, random_actions
    envs = []
    for _ in range(2):
        env, co, E, state, info = make_pair(N=N, K=1, mode="legacy", use_hidden=True, seed=17, obs_mode="rgb_u8")
        envs.append(env)
    envs[0]._can_fuse_render = lambda: False
    rng = np.random.default_rng(8)
    for step in range(6):
        act = random_actions(rng, N, shoot_p=0.9)
        a = envs[0].stateless_step(act)
        b = envs[1].stateless_step(act)
        assert torch.equal(a[0][0], b[0][0]), step


@pytest.mark.gpu
@pytest.mark.parametrize("tag,mode", [("ext", "rgb_f32"), ("plain", "rgb_f32"), ("ext", "rgb_u8")])
def test_reset_frame_reference_mode(cuda_device, tag, mode):
    """reset_obs="reference": reset() returns the frame the reference's reset() returns for the same initial sample
    (reference-source golden, section reset_frame); the default shows the true grid, which every later step shows."""
    import torch
    from gym_cellular_automata_b200.forest_fire.bulldozer.advanced_bulldozer import AdvancedForestFireBulldozerEnv
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_shim_golden.npz"))
    g = z[f"reset_frame/{tag}/sample"].astype(np.float32)
    pos, rgb = z[f"reset_frame/{tag}/position"], z[f"reset_frame/{tag}/rgb"]
    N, H, W, _ = g.shape
    frames = {}
    for how in ("reference", "true_grid"):
        env = AdvancedForestFireBulldozerEnv(H, W, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=False,
                                             enable_extensions=(tag == "ext"), obs_mode=mode, reset_obs=how, seed=3)
        env._ensure_initial()
        assert env._init_grid5.shape == g.shape
        env._init_grid5 = g.copy()
        env._init_context["per_env_context"]["true_grid"] = np.ascontiguousarray(g[..., 0])
        env._init_context["position"] = pos.copy()
        (frame, _ctx), _info = env.reset()
        frames[how] = frame.cpu().numpy()
    want = rgb.astype(np.uint8) if mode == "rgb_u8" else rgb
    assert frames["reference"].dtype == want.dtype and np.array_equal(frames["reference"], want)
    assert not np.array_equal(frames["true_grid"], want)


@pytest.mark.gpu
@pytest.mark.parametrize("K", [3, 4])
def test_tiled_fused_auto_reset_matches_two_call_path(cuda_device, K):
    """The tiled path (a 96 x 80 grid: 3 x 2 tiles, the last ones partial; odd K ends in the scratch grid and is copied
    back) with the fused conditional_reset as a node of the step's CUDA graph == stateless_step followed by
    conditional_reset, state by state, while envs terminate and restart."""
    from parity_util import make_pair, random_actions, sync, read_cuda_state
    N = 5
    envs = []
    for fused in (False, True):
        env, co, E, state, info = make_pair(N=N, size=96, ncols=80, K=K, mode="legacy", use_hidden=True, seed=21,
                                            fast_slope=True)
        ctx = state["per_env_context"]
        ctx["fire_age"][ctx["true_grid"] == 2] = np.float32(3)
        sync(env, state, as_snapshot=True)
        envs.append(env)
    rng = np.random.default_rng(6)
    seen_done = 0
    for step in range(10):
        act = random_actions(rng, N)
        a = torch.as_tensor(act, device="cuda")
        tup = envs[0].stateless_step(act)
        seen_done += int(tup[2].sum())
        tup = envs[0].conditional_reset(tup, act)
        out = envs[1].step_device(a, auto_reset=True)
        s0, s1 = read_cuda_state(envs[0]), read_cuda_state(envs[1])
        for k in s0:
            assert np.array_equal(s0[k], s1[k]), (step, k)
        assert np.array_equal(tup[1].cpu().numpy(), out.reward.cpu().numpy())
    assert seen_done > 0
