/* switchfl_b200.h -- C-ABI of the B200 backend for SwitchFL's lockstep hot path.
 *
 * The reference (AI4REALNET/network-distributed-q-learning) is pure Python and has no FFI of its own;
 * the seam this library sits behind is the pair of Python classes used by main.py / test_model.py /
 * hyperparam_tuning.py.  Each entry point names the reference interface it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions: every function returns 0 on success or a negative SFL_E_* code and never throws; all
 * bulk buffers are CALLER-OWNED device pointers (the Python host allocates them as torch tensors) and
 * the library never frees them; only the small map-constant block is owned by the context.  One
 * context per GPU per map, driven from one host thread; work is enqueued on the caller's stream
 * (a cudaStream_t passed as void*).  Several contexts (maps, devices) may live in one process and run on different streams at
 * the same time: a launch carries its map / layout / arguments as kernel parameters, there is no process-global device
 * state.  Every entry point runs on the device of its context and restores the caller's current device.  There is no CPU
 * path: without a CUDA device sfl_create fails.
 */
#ifndef SWITCHFL_B200_H
#define SWITCHFL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFL_ABI_VERSION 6

enum {
  SFL_OK = 0,
  SFL_E_ARG = -1,       /* bad argument / unsupported size                                   */
  SFL_E_CUDA = -2,      /* CUDA runtime error (see sfl_last_error)                           */
  SFL_E_NOMEM = -3,     /* caller buffer too small                                           */
  SFL_E_STATE = -4      /* call order (e.g. run before bind)                                 */
};

/* per-environment error bits, reported in sfl_env_counters.err (the reference raises instead)       */
enum {
  SFL_ERR_NO_TRAIN_AT_SWITCH = 1,   /* observer.py:294-307: the reference logs "Bug detected" and then dies on an unbound
                                       local (:307); here the episode is abandoned (counted in `aborted`) and the env resets */
  SFL_ERR_INF_DISTANCE = 2,         /* observer.py:35-36  ValueError                                 */
  SFL_ERR_Q_FULL = 4,               /* per-env Q hash table full                                     */
  SFL_ERR_PEND_FULL = 8,            /* pending-update list of a train full (distr_q.py:340-342)      */
  SFL_ERR_PLAN_FULL = 16,           /* train_action_plan longer than 4                               */
  SFL_ERR_REPLAY_UNDERRUN = 32,     /* replay action stream exhausted                                */
  SFL_ERR_BAD_ACTION = 64,          /* switch_env.py:213-215 assertion                               */
  SFL_ERR_REPLAY_DIVERGED = 128     /* replay: an action recorded as greedy is not the argmax here   */
};

/* run modes (sfl_run) */
enum {
  SFL_MODE_LEARN = 0,    /* distr_q.py:296-366  epsilon-greedy + Q-update (Philox instead of PCG64)  */
  SFL_MODE_GREEDY = 1,   /* distr_q.py:195-224  test(): max_action only, no update                   */
  SFL_MODE_REPLAY = 2,   /* learn() with the actions (and malfunction events) of a recorded trace; an action
                            with bit 6 set was an exploit choice: the row is looked up (inserted) and the
                            recorded action must equal the argmax                                     */
  SFL_MODE_STEP = 3      /* host-driven AEC protocol (switch_env.py:616-666): one launch applies the action the
                            host chose for the waiting decision (replay_act[env * act_cap]), advances the trains to
                            the next decision point and reports it in step_out; no learning on the device            */
};

/* Map + line + timetable constants, all HOST pointers; copied by sfl_create.
 * Replaces the construction work of ASyncSwitchEnv.__init__ -> RailNetwork.__init__
 * (switch_env.py:30-52, rail_network.py:84-133) and the per-reset constants of switch_env.py:93-158;
 * the tables are produced by railmap.py (rows A0, F6, Q4 of SURVEY.md section 8a).                   */
typedef struct sfl_map_desc {
  int32_t H, W;                 /* grid size (square: rail_graph.py:43-48)                            */
  int32_t S, NP, NA, T, NT;     /* switches, ports, moving actions (sum of A-1), trains, targets      */
  int32_t max_episode_steps;    /* flatland timetable (flatland_patch/timetable_generators.py:94-101) */
  const uint16_t *grid;         /* [H*W] flatland 16-bit transitions                                  */
  const int32_t *cell_switch;   /* [H*W] switch index or -1                                           */
  const int32_t *sw_P, *sw_A, *sw_port0, *sw_act0;   /* [S], [S], [S+1], [S+1]                        */
  const int32_t *port_nbr, *port_dist, *port_n_intra, *port_intra0;   /* [NP]                         */
  const int32_t *act_in, *act_out, *act_move;        /* [NA] local port indices, RailEnvActions       */
  const int32_t *init_cell, *init_dir, *target_cell, *ed, *la;        /* [T]                          */
  const int32_t *first_port, *first_dist, *init_delay, *tgt_index;    /* [T] (switch_env.py:507-568)  */
  const int32_t *dist;          /* [NT*H*W*4] distance map, 0x3FFFFFFF = unreachable                  */
  const int8_t *qinit_act;      /* [NP*NT] optimistic action or -1 (distr_q.py:81-181)                */
  const uint8_t *qinit_final;   /* [NP*NT] 1 -> 1000. (last leg), 0 -> 500.                           */
} sfl_map_desc;

typedef struct sfl_config {
  int32_t n_envs;
  int32_t q_cap;                /* Q hash rows per environment, power of two                          */
  int32_t pend_cap;             /* pending updates per train (update_dict, distr_q.py:283)            */
  int32_t max_steps;            /* ASyncSwitchEnv max_steps (100000 in every script)                  */
  int32_t dec_cap, tick_cap;    /* trace capacities per env (0 = tracing off)                         */
  int32_t ep_cap;               /* episode-log capacity per env                                       */
  int32_t act_cap, ev_cap;      /* replay capacities per env                                          */
  int32_t trace_sem;            /* 1: also log the semaphore table after each decision                */
  int32_t shared_q;             /* 1: shared-table mode (extension, no counterpart in the reference): every environment
                                   of this context reads ONE dense Q table and accumulates its TD steps into a delta
                                   buffer; sfl_shared_q_apply folds the mean step into the table  */
} sfl_config;

/* DistrQLearning.__init__ arguments (distr_q.py:32) + MalfunctionParameters (main.py:28-33), per env */
typedef struct sfl_hparams {
  double gamma, epsilon, epsilon_decay_rate, lr, lr_decay_rate, default_q;
  uint64_t seed;                /* Philox key                                                         */
  uint32_t malf_threshold;      /* floor((1 - exp(-rate)) * 2^32), 0 = no malfunctions                */
  int32_t malf_min, malf_max;   /* duration = min + U{0..max-min} + 1                                 */
  int32_t episodes;             /* halt after this many episodes since sfl_reset (<0: never)          */
  int32_t episode_base;         /* global index of the first episode after sfl_reset (RNG stream offset) */
  uint32_t malf_thr2;           /* floor(malf_threshold * 256 / ceil(malf_threshold / 2^24)): stage-2 threshold of the
                                   two-stage malfunction draw; 0 when malf_threshold is 0  */
} sfl_hparams;

/* byte sizes of the caller-owned buffers for a (map, config) pair */
typedef struct sfl_sizes {
  uint64_t state_bytes;         /* env state incl. Q tables                                           */
  uint64_t env_stride;          /* bytes per env inside state                                         */
  uint64_t hparams_bytes;       /* n_envs * sizeof(sfl_hparams)                                       */
  uint64_t trace_dec_bytes, trace_tick_bytes, trace_sem_bytes;
  uint64_t ep_log_bytes, ep_delay_bytes;
  uint64_t replay_act_bytes, replay_ev_bytes;
  uint64_t counters_bytes;      /* n_envs * sizeof(sfl_env_counters)                                  */
  uint64_t step_out_bytes;      /* n_envs * sizeof(sfl_step_rec) (SFL_MODE_STEP only)                 */
  uint64_t shared_q_bytes;      /* shared-table mode: NP*NT*48 rows x a_max doubles (the table)       */
  uint64_t shared_d_bytes;      /*   ... x a_max int64 (sum of TD steps, fixed point 2^-24)           */
  uint64_t shared_c_bytes;      /*   ... x a_max int32 (number of TD steps)                           */
  int32_t q_stride;             /* doubles per Q row (1 key slot + A_max)                             */
  int32_t a_max;
} sfl_sizes;

typedef struct sfl_buffers {    /* all device pointers; optional ones may be NULL                     */
  void *state; void *hparams; void *counters;
  void *trace_dec; void *trace_tick; void *trace_sem;
  void *ep_log; void *ep_delay;
  void *replay_act; void *replay_ev;
  void *step_out;
  void *shared_q; void *shared_d; void *shared_c;    /* shared-table mode; shared_d / shared_c may be all-reduced (sum) by the caller */
} sfl_buffers;

typedef struct sfl_env_counters {      /* written by sfl_run for every env                            */
  uint64_t decisions, ticks, train_ticks;   /* lifetime totals                                        */
  int32_t episodes, err, q_rows, halted;
  int32_t n_dec_logged, n_tick_logged, n_ep_logged, elapsed;
  int32_t aborted, reserved;                /* episodes abandoned on SFL_ERR_NO_TRAIN_AT_SWITCH      */
  uint64_t forced_stops;                    /* decisions whose action mask allowed STOP only (switch_agents.py:104-134)   */
  uint64_t stop_actions;                    /* decisions that chose STOP                                                   */
  uint64_t arrived_trains;                  /* trains at their destination, summed over the finished episodes (distr_q.py:364) */
  uint64_t reserved2;
  uint64_t phase_cycles[6];                 /* sfl_set_phase_clock: SM cycles spent per phase -- 0 train ticks (rail_env.step + _move_trains
                                               bookkeeping), 1 observe (last()), 2 action selection, 3 apply + reward (rest of step()),
                                               4 Q-update, 5 episode reset: the split behind the timers main.py:72-78 prints     */
} sfl_env_counters;

typedef struct sfl_dec_rec {           /* one switch-agent decision (trace)                           */
  int32_t ep, tick, sw, train;
  uint32_t key;                        /* dense state index                           */
  int32_t mask, action, next_sw, reward, done;
  uint64_t arrived;
} sfl_dec_rec;

/* SFL_MODE_STEP output per environment: what AECEnv.last() returns for the waiting decision (observer.py:246-308)
 * plus what the previous ASyncSwitchEnv.step() returned (switch_env.py:663-666)                              */
typedef struct sfl_step_rec {
  int32_t pending;                     /* 1: (sw, train) waits for an action; 0: episode over (see done)             */
  int32_t sw, train;
  uint32_t key;                        /* dense state index of the observation                                        */
  int32_t mask;                        /* action mask, bit a = action a allowed                                       */
  int32_t done;                        /* terminated | truncated << 1                                                 */
  int32_t elapsed;                     /* rail_env._elapsed_steps                                                     */
  int32_t last_next_sw;                /* "next_switch" of the decision just applied, -1 if none                      */
  uint64_t arrived;                    /* "arrived_trains" as a bit set                                               */
  int32_t rewards[64];                 /* _cumulative_rewards[sw][train] for every train (switch_env.py:130, 289)     */
} sfl_step_rec;

typedef struct sfl_tick_rec { int32_t pos; int8_t dir, state; int16_t malf; } sfl_tick_rec;
typedef struct sfl_ep_rec {            /* one finished episode (distr_q.py:360-366)                   */
  double cum_reward; int32_t decisions, arrived, num_malfunctions, ticks;
  uint64_t arrived_mask;               /* the trains at their destination ("arrived_trains", :345, 371) */
} sfl_ep_rec;

/* Distance map (flatland DistanceMap as vendored in flatland_patch/distance_map.py:62-167):
 * dist[k][r][c][heading] = moves from (cell, heading) to target_cells[k], 0x3FFFFFFF if unreachable; all four headings
 * of a target cell are 0.  Host pointers in and out (one-off map preprocessing); computed on `device` by one CTA per
 * target relaxing d(cell, o) = 1 + min over the exits h of (cell, o) of d(cell + step(h), h) to its fixed point.     */
int sfl_distance_map(const uint16_t *grid, int32_t H, int32_t W, const int32_t *target_cells, int32_t n_targets,
                     int32_t *dist, int device);

/* Known-answer hook: the library's Q-update arithmetic -- distr_q.py:441-447, fp64, the
 * reference's operator order, no FMA contraction -- evaluated ON THE DEVICE for n rows of host operands
 * {q, lr, reward, gamma, max_next, bootstrap (0: the train stayed at the same switch, :444-447)}; out[i] = new Q(s, a). */
int sfl_kat_q_update(const double *operands, int32_t n, double *out, int device);

int sfl_abi_version(void);
const char *sfl_last_error(void);

/* Sizes of everything the caller must allocate.  No counterpart in the reference, where the Python objects own their
 * memory (env: switch_env.py:42-73, learner: distr_q.py:33-57); here PyTorch does, and the library is told how much.  */
int sfl_query_sizes(const sfl_map_desc *map, const sfl_config *cfg, sfl_sizes *out);

/* replaces ASyncSwitchEnv(rail_env, ...) + DistrQLearning(env, ...)   (main.py:51-60) */
int sfl_create(const sfl_map_desc *map, const sfl_config *cfg, int device, void **ctx);
int sfl_destroy(void *ctx);                       /* env.close() (switch_env.py, called at distr_q.py:241, 379) + garbage collection */
int sfl_bind(void *ctx, const sfl_buffers *bufs); /* hand over the caller-owned device buffers sized by sfl_query_sizes            */

/* zero state + mark every env "needs reset"; the first sfl_run performs env.reset (switch_env.py:93-158)
 * on the device.  keep_q != 0 keeps Q tables and interaction counters (a new learn()/test() call).     */
int sfl_reset(void *ctx, int keep_q, void *stream);

/* Lanes of a warp that cooperate on one environment (1, 2, 4, 8, 16 or 32; 32/lanes environments share a warp and
 * one instruction stream).  sfl_create picks max(lanes that cover the trains in one pass, lanes that still give the
 * batch ~3.5 warps per SM scheduler); this overrides it.  A scheduling choice only: results do not depend on it.  */
int sfl_set_lanes(void *ctx, int lanes);
int sfl_get_lanes(void *ctx);
/* Warps per CTA of the hot-path kernel: 0 = automatic (4, halved while the launch has fewer than 8 CTAs per SM), or 1, 2,
 * 4.  A scheduling choice only.                                                                        */
int sfl_set_cta_warps(void *ctx, int warps);

/* Compile-time variant of the large-map learn kernel: 1 = the one built for 4 CTAs per SM (more registers), 0 = the one
 * for 7, -1 = automatic (roomy when the launch fits one wave anyway).  A scheduling choice only.                   */
int sfl_set_roomy(void *ctx, int roomy);
/* Per-phase cycle counters (sfl_env_counters.phase_cycles) for the wall-clock breakdown main.py:68-78 prints
 * (switch_env.py:67-73 accumulators).  Instrumented runs use the full kernel (the one that also carries the traces);
 * off by default: the production kernels carry no clock reads.                                                       */
int sfl_set_phase_clock(void *ctx, int on);
/* Which kernel instantiation and launch configuration sfl_run(mode) would use now (traced != 0: with decision / tick
 * traces), as text: "k_run<G=..,KIND=..,TH=..,SQ=..,ONE=..,ROOMY=..> grid=.. block=.. smem=..".  No counterpart in the
 * reference; it lets the parity tests name -- and force, with sfl_set_lanes / sfl_set_roomy -- exactly the
 * instantiations the benchmark launches.  Needs no bound buffers.                                                  */
int sfl_describe_launch(void *ctx, int mode, int traced, char *buf, int cap);

/* enable the optimistic initialisation of distr_q.py:299-300 for rows created from now on           */
int sfl_enable_q_init(void *ctx, int on);

/* __init_q_table runs at the first episode of EVERY learn() call and ASSIGNS its rows (distr_q.py:299-300, 156-158, 179-181):
 * overwrite, in every environment's table, the rows of optimistic-init states that already exist (an earlier learn(),
 * load() or test() created them) with their initial values.  Device-side; uses the bound hparams (default_q).            */
int sfl_reapply_q_init(void *ctx, void *stream);

/* advance every env by up to max_ticks flatland ticks (all decisions in between included);
 * replaces the loop body of DistrQLearning.learn / test  (distr_q.py:302-362, 199-224)               */
int sfl_run(void *ctx, int mode, int max_ticks, void *stream);

/* Shared-table mode: Q[s][a] += (sum of the TD steps proposed for (s, a) since the last apply) / (their number), then
 * the delta buffers are cleared.  Across GPUs the caller all-reduces shared_d and shared_c (integer sums: order
 * independent, bit-reproducible) before calling this, which leaves identical tables on every rank.                 */
int sfl_shared_q_apply(void *ctx, void *stream);
/* The same on explicit buffers: q_dst = q_src + mean step of (d, cnt), which are cleared.  With two tables and two
 * accumulator pairs bound alternately (sfl_bind is a pointer swap) the all-reduce + apply of step k can run on a second
 * stream while sfl_run of step k+1 reads the other table: see Engine.run_shared in backend.py and DESIGN.md section 6.   */
int sfl_shared_q_apply_to(void *ctx, const void *q_src, void *q_dst, void *d, void *cnt, void *stream);

/* Sum over all envs of the decisions taken (num_iter, distr_q.py:284, 361) and of the flatland ticks
 * (rail_env._elapsed_steps) after the last run: device reduction, 16-byte D2H.                         */
int sfl_total_decisions(void *ctx, uint64_t *decisions, uint64_t *ticks, void *stream);

/* Q-table export for distr_q_model.pkl (distr_q.py:492-523): rows of env `env` as (key, A_max values) */
int sfl_export_q(void *ctx, int env, uint32_t *keys_host, double *vals_host, int cap_rows, int *n_rows, void *stream);
/* Q-table import (distr_q.py:510-527 load) */
int sfl_import_q(void *ctx, int env, const uint32_t *keys_host, const double *vals_host, int n_rows, void *stream);

#ifdef __cplusplus
}
#endif
#endif
