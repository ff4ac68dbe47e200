/* gca.h -- C ABI of libgca.so: B200 (sm_100a) kernels for the environment-step hot path of
 * frasermince/gym-cellular-automata's advanced bulldozer forest-fire environment.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*).  No
 * allocation, no hidden global state, no host synchronisation inside a step call; functions
 * return 0 on success or a negative gca_status and never throw or exit
 * (gca_last_error() gives the text).  All calls are re-entrant per stream.
 *
 * Reference interfaces replaced (paths relative to /root/reference/gym_cellular_automata/):
 *   gca_params_init        PartiallyObservableForestFireJax.__init__  forest_fire/operators/ca_alexandridis_jax.py:54-160
 *                          + time mappings                             forest_fire/bulldozer/advanced_bulldozer.py:238-246,745-777
 *   gca_alexandridis_step  PartiallyObservableForestFireJax.update     forest_fire/operators/ca_alexandridis_jax.py:426-460 (vmapped)
 *   gca_env_step           jax.vmap(MDP.update) + _award + _is_done    forest_fire/bulldozer/advanced_bulldozer.py:332-399,1103-1133
 *                          (RepeatCAJax.update                         forest_fire/operators/repeat_ca_jax.py:34-71,
 *                           MoveModifyJax.update                       forest_fire/operators/move_modify_jax.py:148-157)
 *   gca_env_step_host      the same transition as a host rollout loop calls it (host actions in, host reward /
 *                          terminated out: jax.device_get around stateless_step)   agents/jax_ppo.py:1150-1166
 *   gca_move_modify        MoveJax.update / ModifyJax.update           forest_fire/operators/move_modify_jax.py:39-62,102-114
 *   gca_reward_done        _award / _is_done / count_cells             forest_fire/bulldozer/advanced_bulldozer.py:597-633,941-953
 *   gca_conditional_reset  conditional_reset                           forest_fire/bulldozer/advanced_bulldozer.py:422-518
 *   gca_render_rgb, gca_render_rgb_actions
 *                          build_observation_on_extensions/grid_to_rgb forest_fire/bulldozer/advanced_bulldozer.py:988-1101
 *                          + apply_blur/apply_extensions               forest_fire/bulldozer/utils/extension_utils.py:99-195
 *   gca_pack_state / gca_unpack_state   the float32/int32 context pytree of _initial_context_distribution
 *                                                                       forest_fire/bulldozer/advanced_bulldozer.py:690-743
 *   gca_generate_hidden    init_vegetation / init_density / init_altitude / get_slope   forest_fire/bulldozer/utils/init_utils.py:10-116,166-200
 *   gca_episode_stats_update  step_env_wrapped's statistics         agents/jax_ppo.py:504-655
 *   gca_windy_env_step / gca_windy_pack / gca_windy_unpack   v3 rule set: WindyForestFire.update, RepeatCA, Move / Modify,
 *                          ForestFireBulldozerEnv reward / done        forest_fire/operators/ca_windy.py:41-139, operators/repeat_ca.py:32-45,
 *                                                                       operators/move_modify.py:39-94, bulldozer/bulldozer.py:196-203,393-400
 *   gca_threefry_bits / gca_threefry_split   jax.random.bits / split (third-party jax, unpinned; see oracle/prng.py)
 */
#ifndef GCA_H_
#define GCA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCA_VERSION 105
#define GCA_MAX_R 10        /* burn kernel radius: ceil(log2(size)) - 2; 10 at 4096 */
#define GCA_MAX_K 8         /* CA sub-steps fused into one env step */

typedef enum gca_status {
  GCA_OK = 0,
  GCA_ERR_ARG = -1,         /* null pointer / bad size / bad enum */
  GCA_ERR_UNSUPPORTED = -2, /* shape or option not supported by this build */
  GCA_ERR_CUDA = -3         /* a CUDA runtime call failed; see gca_last_error() */
} gca_status;

/* jax.random stream layouts */
#define GCA_RNG_LEGACY 0        /* jax_threefry_partitionable = False (JAX < 0.5 default) */
#define GCA_RNG_PARTITIONABLE 1 /* jax_threefry_partitionable = True  (JAX >= 0.5 default) */

/* gca_env_step flags */
#define GCA_FLAG_AUTO_RESET 1u  /* fuse conditional_reset for terminated envs into the step */
#define GCA_FLAG_NO_HIDDEN 2u   /* hidden layers off: veg = den = 3, slope factor 1 (pslope may be NULL) */
#define GCA_FLAG_CA_ONLY 4u     /* run only the CA sub-steps (no clock/move/douse/reward bookkeeping) */
#define GCA_FLAG_NO_TMA 8u      /* tiled path: stage tiles with plain loads instead of TMA */
#define GCA_FLAG_HOST_COPY_IN 32u   /* gca_env_step_host: stage the actions through cudaMemcpyAsync even when mapped */
#define GCA_FLAG_HOST_COPY_OUT 64u  /* gca_env_step_host: copy reward / terminated out with cudaMemcpyAsync even when mapped */
#define GCA_FLAG_HOST_COPY 96u      /* both */
#define GCA_FLAG_RENDER 512u         /* render the step's observation into out->rgb in the step kernel's epilogue (64x64 grids,
                                        no extensions): MDP.grid_to_rgb of the NEW grid / position with the PRE-step dousing marks
                                        and day/night (advanced_bulldozer.py:1103-1133); for an env that auto-resets in this step
                                        the frame conditional_reset draws (restored grid / position, post-step dousing marks and
                                        day/night, :462-487).  float32 [N][H][W][3], or uint8 when out->rgb_u8 != 0 */
#define GCA_FLAG_GENERIC_TILES 1024u /* use the generic tiled kernels (one launch per CA sub-step) even for a grid the whole-grid
                                        bit-board kernel takes (W % 64 == 0, H, W <= 256): tests / A-B measurements */
#define GCA_FLAG_HOST_MAPPED 256u    /* gca_env_step_host: the caller guarantees that host_actions, host_reward, host_terminated and
                                        out->host_done are cudaHostAlloc'ed (pinned, mapped, device address == host address under
                                        unified addressing, e.g. torch pin_memory()): skips the per-call pointer queries */
#define GCA_FLAG_HOST_ASYNC 128u     /* gca_env_step_host: return after the launch; the caller waits with gca_host_wait */
#define GCA_FLAG_WORK_CYCLES 16u /* diagnostics: work[e] receives the elapsed SM clock cycles of env e's step instead of the cost estimate */

/* Constants of one environment family (host POD, passed to kernels by value). */
typedef struct gca_params {
  int32_t H, W;             /* grid rows, cols */
  int32_t R;                /* burn kernel radius */
  int32_t K;                /* CA sub-steps per env step (1 = reference parity) */
  int32_t rng_mode;         /* GCA_RNG_* */
  int32_t age_lo;           /* randint(fire_age_min, fire_age_max): lo, span, mult */
  uint32_t age_span, age_mult;
  int32_t day_length;       /* 400 */
  float p_tree;             /* empty -> tree probability (0 in the reference) */
  float p_wind_change;      /* 0.06 */
  float t_any;              /* f32(0.001) */
  float t_move[9];          /* clock cost of each move action (all equal in the reference) */
  float t_shoot[2];
  float onep_veg[8];        /* f32(1) + p_veg[i], i = 0..5 */
  float onep_den[8];
  float winds[8 * 9];       /* wind matrix k (3x3 row-major, centre 0), float32 */
  float dous_border, dous_inner;   /* 5x5 dousing weights: 16 border cells / 9 inner cells */
  float ring_w[GCA_MAX_R + 1];     /* ring_w[k], k = 1..R: burn-kernel weight of Chebyshev ring k;
                                      ring_w[0] = centre weight (== ring_w[1]) */
} gca_params;

/* Packed device-resident state of N independent environments (all pointers device memory).
 *   cell    u8  [N][H][W]   0 empty / 1 tree / 2 fire
 *   death   u16 [N][H][W]   fire cell: low 16 bits of the CA tick at which it burns out
 *                           (fire_age = ((death - tick) & 0xFFFF) + 1); other cells: fire_age itself
 *   hidden  u8  [N][H][W]   vegetation (bits 0-2) | density (bits 3-5); NULL with GCA_FLAG_NO_HIDDEN
 *   doused  u64 [N][H][WW]  dousing_count bit-board, WW = ceil(W/64), bit (c & 63) of word c >> 6
 *   pslope  f32 [N][H][W][8] exp(f32(0.078) * slope) for the 8 neighbours in 3x3 row-major order,
 *                           centre skipped; NULL = all 1.0
 *   row_min u32 [N][H]      earliest burn-out tick among the fire cells of a row (0xFFFFFFFF: none);
 *                           maintained and needed by the 64x64 kernel only
 *   tick    u32 [N]         CA sub-steps executed since the env was (re)set
 */
typedef struct gca_state {
  int32_t N;
  int32_t reserved;
  uint8_t* cell;
  uint16_t* death;
  const uint8_t* hidden;
  uint64_t* doused;
  const float* pslope;
  uint32_t* row_min;
  uint32_t* tick;
  uint32_t* key;        /* [N][2] raw threefry key data */
  int32_t* wind_index;  /* [N] 0..7 */
  int32_t* position;    /* [N][2] (row, col) */
  float* time;          /* [N] clock accumulator */
  int32_t* time_step;   /* [N] */
  int32_t* is_night;    /* [N] */
  float* steps_elapsed;       /* [N] info bookkeeping (advanced_bulldozer.py:390-391) */
  float* reward_accumulated;  /* [N] */
  /* scratch of the tiled path (grids other than 64x64); may be NULL for 64x64 */
  uint8_t* scratch_cell;      /* [N][H][W] second grid buffer (tiles read one, write the other) */
  uint32_t* scratch_u32;      /* N*98 + 16 + 3*T words, T = N * ceil(H/32) * ceil(W/64) tiles: per env the key schedule
                                 of up to GCA_MAX_K sub-steps (12 words each) + tree/fire counts (2); 16 counters of the
                                 per-sub-step active-tile lists; per tile its number of burning cells; two list buffers */
  /* optional load balancing of the 64x64 kernel (one warp per env, so an env step costs what its
   * fire front costs): the kernel writes work[e]; gca_balance_order turns it into order[] */
  uint32_t* work;             /* [N] cost estimate of the last env step, or NULL */
  const int32_t* order;       /* [N] env index handled by warp slot i (a permutation), or NULL */
  /* tree / fire bit-boards u64 [N][H][WW][2] (bit (c & 63) of word pair c >> 6: tree, fire), a second
   * representation of `cell` kept in step with it by gca_pack_state, gca_conditional_reset and the 64x64
   * step kernel, which reads its grid from here (1 KB per env instead of 4 KB); required for 64x64,
   * ignored (may be NULL) by the tiled path */
  uint64_t* bb;
} gca_state;

/* Per-step outputs (device). Any pointer may be NULL to skip that output. */
typedef struct gca_step_out {
  float* reward;        /* [N] reward handed to the agent (after auto-reset: _award(restored grid)) */
  float* step_reward;   /* [N] reward of the transition itself (info["reward"]) */
  uint8_t* terminated;  /* [N] done flag of the transition (before any auto-reset) */
  int32_t* counts;      /* [N][2] tree, fire cell counts of the post-step grid */
  uint8_t* obs_night;   /* [N] is_night value the observation must be rendered with (pre-flip) */
  unsigned long long* stats; /* [8] += front cells, (cell,dir) draws, ignitions, burn-outs,
                                       threshold cells (exact re-evaluations), env steps, 0, 0 */
  float* host_reward;       /* optional [N] mirror of `reward` in MAPPED pinned host memory (device-visible */
  uint8_t* host_terminated; /* address): the 64x64 step kernel stores both there as well; NULL = off      */
  /* completion word (64x64 kernel, optional): the env whose step ends last stores done_token to *host_done (one
   * word of MAPPED pinned host memory, device-visible address) after every host mirror of the launch is visible
   * to the host -- a host thread polls it (gca_host_wait) instead of synchronising the stream.  done_counter is
   * one zero-initialised device word the launch counts its envs in (it is back to zero when the word is stored). */
  uint32_t* host_done;
  uint32_t* done_counter;
  uint32_t done_token;
  uint32_t rgb_u8;          /* GCA_FLAG_RENDER: 0 = float32 pixels (the reference's layout), else uint8 */
  void* rgb;                /* GCA_FLAG_RENDER: [N][H][W][3] observation of this step (device) */
} gca_step_out;

/* Injected random fields for rule-parity tests (device, each with a leading K axis); NULL = threefry. */
typedef struct gca_inject {
  const float* u_burn;      /* [K][N][H][W][9] */
  const float* u_grow;      /* [K][N][H][W]    */
  const int32_t* age_new;   /* [K][N][H][W]    */
  const float* u_wind;      /* [K][N]          */
  const int32_t* wind_step; /* [K][N]          */
} gca_inject;

int gca_version(void);
const char* gca_last_error(void);
/* Hash of the sources this binary was built from (gym_cellular_automata_b200/_lib.py passes it to nvcc and refuses a
 * library whose id differs from the sources next to it); "unknown" for a build that did not set it. */
const char* gca_build_id(void);

/* A0: fill *p for an nrows x ncols env.  winds72 may be NULL (the reference's 8 wind matrices are
 * generated).  t_move / t_shoot < 0 selects the reference formula from speed_move / speed_act. */
int gca_params_init(gca_params* p, int32_t nrows, int32_t ncols, int32_t K, double speed_move,
                    double speed_act, double t_any, double t_move, double t_shoot, double p_tree,
                    double p_wind_change, int32_t rng_mode, const float* winds72);

/* One full env step for all N envs (clock, K CA sub-steps, move, douse, time_step, day/night,
 * reward, done, info counters, optional fused auto-reset from `snapshot`).  actions: [N][3] int32
 * (move 0-8, shoot 0-1, extension id).  snapshot / snapshot_reward are only read with
 * GCA_FLAG_AUTO_RESET.  State is updated IN PLACE. */
int gca_env_step(const gca_params* p, const gca_state* s, const int32_t* actions,
                 const gca_step_out* out, const gca_inject* inj, const gca_state* snapshot,
                 const float* snapshot_reward, uint32_t flags, void* stream);

/* The same step for a caller whose actions and results live in HOST memory, i.e. the call a CPU-side
 * rollout loop makes once per step; the results are valid on return (the stream is synchronised).
 * Two transports, chosen per call and per direction from the buffers themselves (cudaPointerGetAttributes):
 *  - zero-copy (64x64 grids, host buffer pinned and mapped into the device address space, e.g.
 *    cudaHostAlloc / torch's pin_memory()): the step kernel reads host_actions [N][3] over the bus
 *    itself (dev_actions is not touched) / stores reward and terminated to host_reward [N] and
 *    host_terminated [N] next to its device outputs -- no copy engine in the step;
 *  - staged (any other case, or GCA_FLAG_HOST_COPY_IN / _OUT): host_actions is copied to the device
 *    buffer dev_actions before gca_env_step / out->reward and out->terminated (both required) are copied
 *    out after it (one copy when, on both sides, terminated starts right after the N rewards). */
int gca_env_step_host(const gca_params* p, const gca_state* s, const int32_t* host_actions,
                      int32_t* dev_actions, const gca_step_out* out, const gca_state* snapshot,
                      const float* snapshot_reward, uint32_t flags, float* host_reward,
                      uint8_t* host_terminated, void* stream);

/* Completion of a gca_env_step_host call made with GCA_FLAG_HOST_ASYNC (an EnvPool-style rollout loop steps one
 * group of envs while it handles the results of another): when the call used the completion word
 * (out->host_done / done_counter set, zero-copy transport out) poll host_done -- the HOST address of that word --
 * until it holds `token`; otherwise (host_done NULL) synchronise `stream`.  Returns GCA_ERR_CUDA if the word has
 * not arrived after timeout_s seconds (the stream's error state is then reported by gca_last_error). */
int gca_host_wait(const uint32_t* host_done, uint32_t token, double timeout_s, void* stream);

/* K CA sub-steps only (PartiallyObservableForestFireJax.update applied K times). */
int gca_alexandridis_step(const gca_params* p, const gca_state* s, const gca_step_out* out,
                          const gca_inject* inj, uint32_t flags, void* stream);

/* MoveJax + ModifyJax on positions / dousing bit-board only. */
int gca_move_modify(const gca_params* p, const gca_state* s, const int32_t* actions, void* stream);

/* count_cells + _award + _is_done on the packed grid. */
int gca_reward_done(const gca_params* p, const gca_state* s, float* reward, uint8_t* terminated,
                    int32_t* counts, void* stream);

/* conditional_reset: restore envs with terminated[e] != 0 from `snapshot` (time_step / is_night
 * kept), zero their info counters, set reward[e] = snapshot_reward[e] and clear terminated[e]. */
int gca_conditional_reset(const gca_params* p, const gca_state* s, const gca_state* snapshot,
                          const float* snapshot_reward, float* reward, uint8_t* terminated,
                          void* stream);

/* Observation: float32 RGB [N][H][W][3] (reference layout) or uint8 RGB when rgb_u8 != 0.
 * cell: grid to draw; night: [N] is_night to use; ext_action: [N] extension id (0..2) or NULL;
 * scratch: [N] uint32 device words, required when enable_extensions != 0;
 * env_mask: [N] or NULL -- when given, only envs with env_mask[e] != 0 are (re)drawn, which is how
 * conditional_reset re-renders terminated envs without a host round trip. */
int gca_render_rgb(const gca_params* p, int32_t N, const uint8_t* cell, const uint64_t* doused,
                   const int32_t* position, const uint8_t* night, const int32_t* ext_action,
                   const uint8_t* env_mask, int32_t enable_extensions, int32_t rgb_u8,
                   uint32_t* scratch, void* rgb_out, void* stream);

/* The same with the extension id taken from the step's action triples: actions [N][3] int32 (move, shoot,
 * extension id), i.e. the array handed to gca_env_step -- no gather of the third column needed. */
int gca_render_rgb_actions(const gca_params* p, int32_t N, const uint8_t* cell, const uint64_t* doused,
                           const int32_t* position, const uint8_t* night, const int32_t* actions,
                           const uint8_t* env_mask, int32_t enable_extensions, int32_t rgb_u8,
                           uint32_t* scratch, void* rgb_out, void* stream);

/* Reference float32/int32 context arrays <-> packed state (device pointers).  fire_age of fire
 * cells must be an integer in [1, 32767]; vegetation / density in [0, 7].  err_flag (device int32,
 * may be NULL) is set non-zero if an input violates that. */
int gca_pack_state(const gca_params* p, const gca_state* s, const float* true_grid,
                   const float* fire_age, const int32_t* dousing_count, const int32_t* vegetation,
                   const int32_t* density, uint8_t* hidden_out, int32_t* err_flag, void* stream);
int gca_unpack_state(const gca_params* p, const gca_state* s, float* true_grid, float* fire_age,
                     int32_t* dousing_count, void* stream);

/* Deal the envs to warp slots so that every SM (and scheduler) gets the same share of heavy and
 * light envs: sort by work[] (descending) and distribute round-robin over waves of 148 CTAs in
 * alternating direction.  order and work are [N] device arrays; call it every few steps. */
int gca_balance_order(int32_t N, const uint32_t* work, int32_t* order, void* stream);

/* ---- device-side generation of the hidden layers (inputs of the path) ------------------------------------
 * vegetation / density [N][H][W] i32 (random rectangles of type 1..5 over a 1..3 background), altitude [N][H][W]
 * f32 (noise + cosine hills + ramps, / 10), slope9 [N][H][W][3][3] f32 (degrees, centre 0, border cells flat;
 * may be NULL) and pslope9 = exp(f32(0.078) * slope) in the same layout -- the layer models of
 * forest_fire/bulldozer/utils/init_utils.py:10-116,166-200 and ca_alexandridis_jax.py:199-200, drawn from a
 * counter-based generator: env e of the call gets the layers of global env (env_offset + e) under `seed`,
 * whatever the batch size or sharding.  altitude_f64 is scratch, [N][H][W] doubles.  H, W >= 16. */
int gca_generate_hidden(int32_t N, int32_t H, int32_t W, uint64_t seed, int32_t env_offset, int32_t* vegetation,
                        int32_t* density, float* altitude, double* altitude_f64, float* slope9, float* pslope9,
                        void* stream);

/* ---- rollout-side episode statistics (the caller of the hot path) ----------------------------------------
 * One call = the statistics part of step_env_wrapped (agents/jax_ppo.py:504-655) for all N envs, to be
 * issued after gca_env_step and before any conditional reset is observed by the caller:
 *   episode_returns += step_reward; episode_lengths += 1; day/night extension-correctness counters
 *   (jax_ppo.py:524-541; they are never cleared, like the reference); finished = terminated + truncated;
 *   returned_episode_* latch the totals of finished envs; episode_* are zeroed for them;
 *   amount_finished += sum(terminated); the 10-entry ring buffers receive the finished envs in env
 *   order starting at recent_idx (jax_ppo.py:543-621, a serial lax.scan there; here a block scan), and
 *   recent_idx = (recent_idx + #finished) % 10.
 * All pointers are device memory; per-env arrays have N entries, ring arrays 10, scalars 1. */
#define GCA_RECENT 10
typedef struct gca_episode_stats {
  float* episode_returns;            /* [N] */
  int32_t* episode_lengths;          /* [N] */
  float* returned_episode_returns;   /* [N] */
  int32_t* returned_episode_lengths; /* [N] */
  int32_t* amount_finished;          /* [1] */
  float* recent_returns;             /* [10] */
  int32_t* recent_lengths;           /* [10] */
  int32_t* recent_idx;               /* [1] */
  int32_t* current_day_correct;      /* [N] */
  int32_t* current_night_correct;    /* [N] */
  int32_t* current_day_steps;        /* [N] */
  int32_t* current_night_steps;      /* [N] */
  int32_t* recent_day_correct;       /* [10] */
  int32_t* recent_night_correct;     /* [10] */
  int32_t* recent_day_steps;         /* [10] */
  int32_t* recent_night_steps;       /* [10] */
} gca_episode_stats;
/* step_reward: info["reward"] of the transition (gca_step_out.step_reward); terminated: gca_step_out.terminated;
 * truncated: [N] or NULL (all zero); obs_night: is_night of the observation the action was chosen on
 * (gca_step_out.obs_night); actions: [N][3], the extension id is actions[.][2]. */
int gca_episode_stats_update(int32_t N, const gca_episode_stats* st, const float* step_reward,
                             const uint8_t* terminated, const uint8_t* truncated, const uint8_t* obs_night,
                             const int32_t* actions, void* stream);

/* ---- v3 rule set: WindyForestFire + Move / Modify(cut) / RepeatCA of ForestFireBulldozer256x256-v3 ----
 * (reference forest_fire/operators/ca_windy.py:41-139, operators/repeat_ca.py:32-45,
 *  operators/move_modify.py:39-94, bulldozer/bulldozer.py:196-203,393-400).
 * State: tree / fire bit-boards u64 [N][H][ceil(W/64)], position i32 [N][2], time f64 [N].
 * actions i32 [N][2] (move 0-8, shoot 0-1); wind f64 [9] (3x3 row-major, shared by all envs);
 * rolls f64 [N][rmax][9]: the uniform 3x3 roll of each CA update of this step (the reference draws
 * them from an unseeded generator, so they are an input); an env uses its first `repeats` rolls.
 * H * ceil(W/64) * 24 bytes of shared memory per env must fit one SM (grids up to ~ 512 x 512). */
int gca_windy_env_step(int32_t N, int32_t H, int32_t W, uint64_t* tree_bb, uint64_t* fire_bb, int32_t* position,
                       double* time, const int32_t* actions, const double* wind9, const double* rolls,
                       int32_t rmax, double t_move, double t_shoot, double t_any, double* reward,
                       uint8_t* terminated, int32_t* counts, int32_t* repeats_out, void* stream);
/* u8 cell codes (0 empty, 1 tree, 2 fire = the reference's 0 / 3 / 25) <-> bit-boards */
int gca_windy_pack(int32_t N, int32_t H, int32_t W, const uint8_t* cell, uint64_t* tree_bb, uint64_t* fire_bb,
                   void* stream);
int gca_windy_unpack(int32_t N, int32_t H, int32_t W, const uint64_t* tree_bb, const uint64_t* fire_bb,
                     uint8_t* cell, void* stream);

/* Test hooks for the in-kernel PRNG: n words of jax.random.bits(key,(n,)) / split(key, num). */
int gca_threefry_bits(const uint32_t* key2_dev, int64_t n, int32_t rng_mode, uint32_t* out_dev,
                      void* stream);
int gca_threefry_split(const uint32_t* key2_dev, int32_t num, int32_t rng_mode, uint32_t* out_dev,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCA_H_ */
