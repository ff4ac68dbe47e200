"""Reference-side binding of libgca.so -- the stub a maintainer of frasermince/gym-cellular-automata would add as
``gym_cellular_automata/forest_fire/bulldozer/_gca.py`` to step ``AdvancedForestFireBulldozerEnv`` on a B200.

Self-contained on purpose: ctypes + the C ABI of include/gca.h only -- nothing from gym_cellular_automata_b200.
Device buffers are whatever the caller has (JAX arrays: ``unsafe_buffer_pointer()``; torch: ``data_ptr()``); the
small ``Buffers`` helper below uses torch because that is what this image offers.  INTEGRATION.md shows this file
verbatim; tests/test_host_api.py checks every structure here field by field against the header, and
tests/test_gpu_parity.py steps a 64x64 env through this file alone and compares it with the oracle.

Replaces, in the reference: PartiallyObservableForestFireJax.__init__ constants (ca_alexandridis_jax.py:54-160),
the body of stateless_step (advanced_bulldozer.py:332-399: vmap(MDP.update) + _award + _is_done + info), and
conditional_reset (advanced_bulldozer.py:422-518, here GCA_FLAG_AUTO_RESET inside the step).
"""
import ctypes as C
import os

GCA_VERSION = 105
GCA_FLAG_AUTO_RESET, GCA_FLAG_NO_HIDDEN = 1, 2
GCA_RNG_LEGACY, GCA_RNG_PARTITIONABLE = 0, 1   # jax_threefry_partitionable False / True


class gca_params(C.Structure):                   # include/gca.h: struct gca_params (filled by gca_params_init)
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("R", C.c_int32), ("K", C.c_int32), ("rng_mode", C.c_int32),
                ("age_lo", C.c_int32), ("age_span", C.c_uint32), ("age_mult", C.c_uint32), ("day_length", C.c_int32),
                ("p_tree", C.c_float), ("p_wind_change", C.c_float), ("t_any", C.c_float),
                ("t_move", C.c_float * 9), ("t_shoot", C.c_float * 2), ("onep_veg", C.c_float * 8),
                ("onep_den", C.c_float * 8), ("winds", C.c_float * 72), ("dous_border", C.c_float),
                ("dous_inner", C.c_float), ("ring_w", C.c_float * 11)]


class gca_state(C.Structure):                    # device pointers; shapes in include/gca.h
    _fields_ = [("N", C.c_int32), ("reserved", C.c_int32)] + [(n, C.c_void_p) for n in (
        "cell", "death", "hidden", "doused", "pslope", "row_min", "tick", "key", "wind_index", "position",
        "time", "time_step", "is_night", "steps_elapsed", "reward_accumulated", "scratch_cell", "scratch_u32",
        "work", "order", "bb")]


class gca_step_out(C.Structure):                 # per-step outputs; any pointer may be NULL
    _fields_ = [(n, C.c_void_p) for n in ("reward", "step_reward", "terminated", "counts", "obs_night", "stats",
                                           "host_reward", "host_terminated", "host_done", "done_counter")] + [
        ("done_token", C.c_uint32), ("rgb_u8", C.c_uint32), ("rgb", C.c_void_p)]


class gca_inject(C.Structure):                   # injected random fields (rule-parity tests); NULL = in-kernel threefry
    _fields_ = [(n, C.c_void_p) for n in ("u_burn", "u_grow", "age_new", "u_wind", "wind_step")]


def load(path=None):
    lib = C.CDLL(path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gym_cellular_automata_b200",
                                      "libgca.so"))   # built by: python -c "import __graft_entry__ as g; g.build()"
    lib.gca_last_error.restype = C.c_char_p
    if lib.gca_version() != GCA_VERSION:
        raise RuntimeError(f"libgca.so is version {lib.gca_version()}, this binding expects {GCA_VERSION}")
    lib.gca_params_init.argtypes = [C.POINTER(gca_params), C.c_int32, C.c_int32, C.c_int32] + [C.c_double] * 7 + [
        C.c_int32, C.c_void_p]
    lib.gca_pack_state.argtypes = [C.POINTER(gca_params), C.POINTER(gca_state)] + [C.c_void_p] * 8
    lib.gca_unpack_state.argtypes = [C.POINTER(gca_params), C.POINTER(gca_state)] + [C.c_void_p] * 4
    lib.gca_reward_done.argtypes = [C.POINTER(gca_params), C.POINTER(gca_state)] + [C.c_void_p] * 4
    lib.gca_env_step.argtypes = [C.POINTER(gca_params), C.POINTER(gca_state), C.c_void_p, C.POINTER(gca_step_out),
                                 C.POINTER(gca_inject), C.POINTER(gca_state), C.c_void_p, C.c_uint32, C.c_void_p]
    lib.gca_render_rgb_actions.argtypes = [C.POINTER(gca_params), C.c_int32] + [C.c_void_p] * 6 + [
        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def check(lib, rc):
    if rc:
        raise RuntimeError(lib.gca_last_error().decode())


class Buffers:
    """Device buffers of N packed 64x64 envs (what a JAX maintainer would allocate with jnp.zeros / device_put)."""

    def __init__(self, N, H, W, use_hidden, torch):
        d, z = "cuda", torch.zeros
        self.t = dict(
            cell=z((N, H, W), dtype=torch.uint8, device=d), death=z((N, H, W), dtype=torch.uint16, device=d),
            hidden=z((N, H, W), dtype=torch.uint8, device=d) if use_hidden else None,
            doused=z((N, H, (W + 63) // 64), dtype=torch.int64, device=d),
            pslope=torch.ones((N, H, W, 8), dtype=torch.float32, device=d) if use_hidden else None,
            row_min=torch.full((N, H), -1, dtype=torch.int32, device=d), tick=z(N, dtype=torch.int32, device=d),
            key=z((N, 2), dtype=torch.uint32, device=d), wind_index=z(N, dtype=torch.int32, device=d),
            position=z((N, 2), dtype=torch.int32, device=d), time=z(N, dtype=torch.float32, device=d),
            time_step=z(N, dtype=torch.int32, device=d), is_night=z(N, dtype=torch.int32, device=d),
            steps_elapsed=z(N, dtype=torch.float32, device=d), reward_accumulated=z(N, dtype=torch.float32, device=d),
            bb=z((N, H, (W + 63) // 64, 2), dtype=torch.int64, device=d))   # tree / fire bit-boards: REQUIRED for 64x64
        self.struct = gca_state(N=N)
        for k, v in self.t.items():
            setattr(self.struct, k, None if v is None else v.data_ptr())

    def clone(self, torch):
        o = Buffers.__new__(Buffers)
        o.t = {k: (v if v is None or k in ("hidden", "pslope") else v.clone()) for k, v in self.t.items()}
        o.struct = gca_state(N=self.struct.N)
        for k, v in o.t.items():
            setattr(o.struct, k, None if v is None else v.data_ptr())
        return o


class RefSideEnv:
    """What the reference's env keeps between calls once its stateless_step body is the one C call below."""

    def __init__(self, lib, torch, nrows, ncols, num_envs, speed_move, speed_act, substeps=1, rng_mode=GCA_RNG_LEGACY,
                 use_hidden=True, auto_reset=True):
        self.lib, self.torch, self.N, self.H, self.W = lib, torch, num_envs, nrows, ncols
        self.params = gca_params()
        # A0: replaces the constants of PartiallyObservableForestFireJax.__init__ and the clock mapping
        check(lib, lib.gca_params_init(C.byref(self.params), nrows, ncols, substeps, speed_move, speed_act, 0.001, -1.0, -1.0,
                                       0.0, 0.06, rng_mode, None))
        self.flags = (0 if use_hidden else GCA_FLAG_NO_HIDDEN) | (GCA_FLAG_AUTO_RESET if auto_reset else 0)
        self.state = Buffers(num_envs, nrows, ncols, use_hidden, torch)
        z = torch.zeros
        self.reward = z(num_envs, dtype=torch.float32, device="cuda")
        self.step_reward = z(num_envs, dtype=torch.float32, device="cuda")
        self.terminated = z(num_envs, dtype=torch.uint8, device="cuda")
        self.counts = z((num_envs, 2), dtype=torch.int32, device="cuda")
        self.obs_night = z(num_envs, dtype=torch.uint8, device="cuda")
        self.out = gca_step_out(reward=self.reward.data_ptr(), step_reward=self.step_reward.data_ptr(),
                                terminated=self.terminated.data_ptr(), counts=self.counts.data_ptr(),
                                obs_night=self.obs_night.data_ptr())
        self.snapshot = None

    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def reset(self, ctx, position, time, pslope8=None):
        """reset(): move the reference's float32 / int32 context pytree (advanced_bulldozer.py:690-743) into the packed
        device state; ``ctx`` holds device tensors true_grid, fire_age (float32), dousing_count, vegetation, density,
        wind_index, is_night, time_step (int32) and key (uint32 (N,2)); ``pslope8`` = exp(f32(0.078) * slope) for the 8
        neighbours (N,H,W,8)."""
        t, s = self.state.t, self.state.struct
        if pslope8 is not None:
            t["pslope"].copy_(pslope8)
        check(self.lib, self.lib.gca_pack_state(
            C.byref(self.params), C.byref(s), ctx["true_grid"].data_ptr(), ctx["fire_age"].data_ptr(),
            ctx["dousing_count"].data_ptr(), None if t["hidden"] is None else ctx["vegetation"].data_ptr(),
            None if t["hidden"] is None else ctx["density"].data_ptr(), s.hidden, None, self.stream()))
        for k in ("wind_index", "is_night", "time_step", "key"):
            t[k].copy_(ctx[k].reshape(t[k].shape))
        t["position"].copy_(position)
        t["time"].copy_(time)
        self.snapshot = self.state.clone(self.torch)      # conditional_reset restores terminated envs from it
        self.snap_reward = self.torch.zeros(self.N, dtype=self.torch.float32, device="cuda")
        check(self.lib, self.lib.gca_reward_done(C.byref(self.params), C.byref(self.snapshot.struct),
                                                 self.snap_reward.data_ptr(), None, None, self.stream()))

    def stateless_step(self, actions):
        """stateless_step(action, obs, info): ONE launch.  ``actions`` (N,3) int32 device tensor [move, shoot, extension id]."""
        check(self.lib, self.lib.gca_env_step(C.byref(self.params), C.byref(self.state.struct), actions.data_ptr(),
                                              C.byref(self.out), None, C.byref(self.snapshot.struct),
                                              self.snap_reward.data_ptr(), self.flags, self.stream()))
        return self.reward, self.terminated

    def observation(self, actions, enable_extensions=False):
        """MDP.grid_to_rgb of the step just made: float32 (N,H,W,3), reference layout."""
        rgb = self.torch.empty((self.N, self.H, self.W, 3), dtype=self.torch.float32, device="cuda")
        scratch = self.torch.zeros(self.N, dtype=self.torch.int32, device="cuda")
        s = self.state.struct
        check(self.lib, self.lib.gca_render_rgb_actions(C.byref(self.params), self.N, s.cell, s.doused, s.position,
                                                        self.obs_night.data_ptr(), actions.data_ptr(), None,
                                                        int(enable_extensions), 0, scratch.data_ptr(), rgb.data_ptr(),
                                                        self.stream()))
        return rgb

    def context(self):
        """The reference's true_grid / fire_age (float32) and dousing_count (int32) of the current state."""
        g = self.torch.empty((self.N, self.H, self.W), dtype=self.torch.float32, device="cuda")
        a = self.torch.empty_like(g)
        d = self.torch.empty((self.N, self.H, self.W), dtype=self.torch.int32, device="cuda")
        check(self.lib, self.lib.gca_unpack_state(C.byref(self.params), C.byref(self.state.struct), g.data_ptr(),
                                                  a.data_ptr(), d.data_ptr(), self.stream()))
        return {"true_grid": g, "fire_age": a, "dousing_count": d}
