// sfl_core.cuh -- per-environment logic of SwitchFL's lockstep hot path, one warp per environment.
//
// Restates (not ports) rows E1-E7, O1-O3, R1, Q1-Q3, F1-F5 of SURVEY.md section 8a for a batched,
// structure-of-arrays device layout.  Citations "file:line" are relative to the reference repository.
//
// Execution model: one warp owns one environment for the whole launch.  The per-tick train phase is
// lane-parallel (lane = train, strided for T > 32) with warp-level reductions; the per-decision phase
// is inherently serial inside an environment (every _apply_action mutates the semaphores the next
// observe reads, switch_env.py:648) and runs on lane 0.  Cross-lane communication goes only through the
// per-warp Scratch block and the reductions w_or64 / w_any, so the same source also compiles for a
// single "lane" on the host: tests/emul builds it with g++ (-DSFL_HOST_EMUL) to unit-test the logic on
// the CPU box that has no GPU.  That build is test infrastructure; the product library has no CPU path.
#pragma once
#include <stdint.h>
#include "switchfl_b200.h"

#if defined(__CUDACC__) && !defined(SFL_HOST_EMUL)
#define SFL_FN __device__ __forceinline__
#define SFL_NI __device__ __noinline__
#define SFL_CONST __constant__
#define SFL_LANES 32
#else
#include <math.h>
#include <string.h>
#define SFL_FN inline
#define SFL_NI inline
#define SFL_CONST static
#define SFL_LANES 1
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
#endif

namespace sfl {

// flatland enums
enum { A_NOTHING = 0, A_LEFT = 1, A_FWD = 2, A_RIGHT = 3, A_STOP = 4, A_NONE = 0xFF };
enum { ST_WAITING = 0, ST_READY = 1, ST_MALF_OFF = 2, ST_MOVING = 3, ST_STOPPED = 4, ST_MALF = 5, ST_DONE = 6 };
enum { SEM_IN = 0, SEM_OUT = 1 };
#define SFL_INF_DIST 0x3FFFFFFF
#define SFL_MAX_T 64
#define SFL_PLAN_CAP 4

// ------------------------------------------------------------------------------------------------ warp primitives
#if SFL_LANES == 32
SFL_FN int lane_id() { return threadIdx.x & 31; }
SFL_FN void w_sync() { __syncwarp(); }
SFL_FN int w_any(int p) { return __any_sync(0xffffffffu, p); }
SFL_FN unsigned long long w_or64(unsigned long long v) {
  unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
  unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
  return ((unsigned long long)hi << 32) | lo;
}
SFL_FN int w_sum(int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
SFL_FN int popc64(unsigned long long v) { return __popcll(v); }
SFL_FN int ffs64(unsigned long long v) { return __ffsll((long long)v) - 1; }
SFL_FN int popc32(unsigned v) { return __popc(v); }
SFL_FN int clz32(unsigned v) { return __clz((int)v); }
SFL_FN double dmul(double a, double b) { return __dmul_rn(a, b); }
SFL_FN double dadd(double a, double b) { return __dadd_rn(a, b); }
#else
SFL_FN int lane_id() { return 0; }
SFL_FN void w_sync() {}
SFL_FN int w_any(int p) { return p; }
SFL_FN unsigned long long w_or64(unsigned long long v) { return v; }
SFL_FN int w_sum(int v) { return v; }
SFL_FN int popc64(unsigned long long v) { return __builtin_popcountll(v); }
SFL_FN int ffs64(unsigned long long v) { return __builtin_ffsll((long long)v) - 1; }
SFL_FN int popc32(unsigned v) { return __builtin_popcount(v); }
SFL_FN int clz32(unsigned v) { return v ? __builtin_clz(v) : 32; }
SFL_FN double dmul(double a, double b) { volatile double r = a * b; return r; }
SFL_FN double dadd(double a, double b) { volatile double r = a + b; return r; }
#endif

// ------------------------------------------------------------------------------------------------ device views
// read-only map table: loads go through the non-coherent path (LDG.CONSTANT, L1-resident across the launch)
template <class T> struct RO {
  const T *p;
#if SFL_LANES == 32
  SFL_FN T operator[](size_t i) const { return __ldg(p + i); }
#else
  SFL_FN T operator[](size_t i) const { return p[i]; }
#endif
};

struct DevMap {            // map constants (device pointers), passed by value to the kernels
  int H, W, Hp, Wp, S, NP, NA, T, NT, max_episode_steps, a_max, pad0;
  RO<uint16_t> grid;           // [Hp*Wp] zero border of 1 cell: moves from a rail cell never leave the array
  RO<int16_t> cell_switch;     // [Hp*Wp]
  RO<int4> sw;                 // [S]  {P, A, port0, act0}
  RO<int4> port;               // [NP] {nbr, dist, n_intra, intra0}
  RO<int16_t> port_switch;     // [NP]
  RO<int4> act;                // [NA] {in_local, out_local, move, 0}
  RO<int4> train0;             // [T]  {init_cell, init_dir, target_cell, tgt_index}   (padded cell ids)
  RO<int4> train1;             // [T]  {ed, la, first_port, first_dist}
  RO<int> init_delay;          // [T]
  RO<int> dist;                // [NT][Hp*Wp][4]
  RO<int8_t> qinit;            // [NP*NT]  -1 | action | final<<4
};

struct Layout {            // byte offsets inside one env block
  int T, S, NP, NT, a_max, q_cap, q_stride, pend_cap;
  unsigned off_pos, off_last_delay, off_malf, off_next_port, off_prev_port, off_source_port, off_act_switch;
  unsigned off_dir, off_state, off_saved, off_prev_act, off_plan_len, off_plan, off_pend_n;
  unsigned off_pend_key, off_pend_meta, off_sem, off_rewards, off_ninter, off_q, pad;
  unsigned long long env_stride;
};

struct EnvHdr {            // 128 bytes at the start of every env block
  int elapsed, episode, step_counter, num_malf;
  int terminated, truncated, need_reset, halted;
  int err, q_rows, pending_fin, cur_dec;
  int act_cursor, ev_cursor, n_dec_logged, n_tick_logged;
  int n_ep_logged, q_init_on, pad0, pad1;
  unsigned long long active_mask, malf_prev_mask, at_dest_mask, done_mask;
  unsigned long long decisions, ticks, train_ticks;
  double cum_reward;
};

struct Env {               // three base pointers; field addresses are base + constant-bank offset
  char *hot;               // staged copy (shared memory) of the first hot_bytes of the env block; == gb on the host build
  char *semb;              // base the semaphore offset applies to (hot when the records are staged, else gb)
  char *gb;                // the env block in HBM
  SFL_FN EnvHdr *h() const { return (EnvHdr *)hot; }
  SFL_FN int *pos() const;          SFL_FN int *last_delay() const;
  SFL_FN int16_t *malf() const;     SFL_FN int16_t *next_port() const;   SFL_FN int16_t *prev_port() const;
  SFL_FN int16_t *source_port() const; SFL_FN int16_t *act_switch() const;
  SFL_FN uint8_t *dir() const;      SFL_FN uint8_t *state() const;       SFL_FN uint8_t *saved() const;
  SFL_FN uint8_t *prev_act() const; SFL_FN uint8_t *plan_len() const;    SFL_FN uint8_t *plan() const;
  SFL_FN uint8_t *pend_n() const;   SFL_FN uint32_t *pend_key() const;   SFL_FN uint32_t *pend_meta() const;
  SFL_FN int4 *sem() const;         // {t0, t1, train (-1 = absent), type}
  SFL_FN int *rewards() const;      SFL_FN int *ninter() const;          SFL_FN double *q() const;
};

struct Scratch {           // per-warp exchange area (shared memory on the device)
  int src[SFL_MAX_T], dst[SFL_MAX_T], ndir[SFL_MAX_T], occ[SFL_MAX_T], exp_cell[SFL_MAX_T], inj[SFL_MAX_T];
  uint8_t pre[SFL_MAX_T], blk[SFL_MAX_T], act[SFL_MAX_T], exp_flags[SFL_MAX_T];
};

struct RunArgs {           // per-launch arguments
  int mode, max_ticks, n_envs, trace_sem;
  int dec_cap, tick_cap, ep_cap, act_cap, ev_cap, max_steps, pad0, pad1;
  char *state;
  const sfl_hparams *hp;
  sfl_env_counters *counters;
  sfl_dec_rec *trace_dec; sfl_tick_rec *trace_tick; int4 *trace_sem_buf;
  sfl_ep_rec *ep_log; int *ep_delay;
  const int8_t *replay_act; const int *replay_ev;     // ev: [env][ev_cap][3] = (tick, train, duration), tick-sorted, tick<0 ends
};

// map constants, env-block layout and launch arguments live in constant memory (set by sfl_run / sfl_reset before the
// launch, in stream order), so the non-inlined device functions below read them as constant-bank operands instead of
// receiving structs by reference.  All contexts of one process must therefore launch on one stream.
SFL_CONST DevMap c_m;
SFL_CONST Layout c_L;
SFL_CONST RunArgs c_ra;

#define SFL_ACC(T, name, base) SFL_FN T *Env::name() const { return (T *)(base + c_L.off_##name); }
SFL_ACC(int, pos, hot) SFL_ACC(int, last_delay, hot) SFL_ACC(int16_t, malf, hot) SFL_ACC(int16_t, next_port, hot)
SFL_ACC(int16_t, prev_port, hot) SFL_ACC(int16_t, source_port, hot) SFL_ACC(int16_t, act_switch, hot)
SFL_ACC(uint8_t, dir, hot) SFL_ACC(uint8_t, state, hot) SFL_ACC(uint8_t, saved, hot) SFL_ACC(uint8_t, prev_act, hot)
SFL_ACC(uint8_t, plan_len, hot) SFL_ACC(uint8_t, plan, hot) SFL_ACC(uint8_t, pend_n, hot)
SFL_ACC(uint32_t, pend_key, hot) SFL_ACC(uint32_t, pend_meta, hot) SFL_ACC(int4, sem, semb)
SFL_ACC(int, rewards, gb) SFL_ACC(int, ninter, gb) SFL_ACC(double, q, gb)
#undef SFL_ACC

// ------------------------------------------------------------------------------------------------ Philox4x32-10
struct U4 { unsigned x, y, z, w; };
SFL_FN unsigned mulhi32(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
SFL_FN U4 philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    unsigned h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    unsigned h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    unsigned n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  U4 r = {c0, c1, c2, c3};
  return r;
}

// ------------------------------------------------------------------------------------------------ F1
struct Chk { int cell, dir, valid, cell_ok; };

// flatland rail.check_action_on_agent; called switch_env.py:325,450,545, reward_func.py:49
SFL_NI Chk check_action(int action, int cell, int dir) {
  unsigned v = c_m.grid[cell];
  unsigned nib = (v >> ((3 - dir) * 4)) & 0xFu;
  int n = popc32(nib);
  int nd = dir, valid = -1;
  if (action == A_LEFT) { nd = dir - 1; if (n <= 1) valid = 0; }
  else if (action == A_RIGHT) { nd = dir + 1; if (n <= 1) valid = 0; }
  nd &= 3;
  if (action == A_FWD && n == 1) { nd = 3 - (31 - clz32(nib)); valid = 1; }
  int delta = (nd == 0) ? -c_m.Wp : (nd == 1) ? 1 : (nd == 2) ? c_m.Wp : -1;
  Chk r;
  r.cell = (v != 0) ? cell + delta : cell;      // a non-rail cell (only reachable by projecting invalid plans) stays put
  r.dir = nd;
  if (valid < 0) valid = (nib >> (3 - nd)) & 1;
  r.valid = valid;
  r.cell_ok = c_m.grid[r.cell] != 0;
  return r;
}

SFL_FN int is_moving(int a) { return a == A_LEFT || a == A_FWD || a == A_RIGHT; }

// ------------------------------------------------------------------------------------------------ E4
// observer.py:44-151 check_port_blocked.  Every writer of a semaphore record stores dir == map_direction(port)
// (rail_network.py:243,327,337,372,382,396,408; switch_env.py:382,566), so the eight clauses reduce to:
//   rule_next: 'out' -> blocked, 'in' -> blocked iff holder MALFUNCTION;  rule_out: 'in' -> blocked, 'out' -> iff MALFUNCTION.
SFL_FN int rule_port(Env e, int port, int me, int now, int blocking_type) {
  int4 r = e.sem()[port];
  if (r.z < 0 || r.z == me || r.x > now || r.y < now) return 0;
  if (r.w == blocking_type) return 1;
  return e.state()[r.z] == ST_MALF;
}
SFL_NI int port_blocked(Env e, int next_port, int out_port, int me, int now) {
  if (next_port >= 0) {
    if (rule_port(e, next_port, me, now, SEM_OUT)) return 1;
    return rule_port(e, out_port, me, now, SEM_IN);
  }
  return rule_port(e, out_port, me, now, SEM_OUT);
}

// ------------------------------------------------------------------------------------------------ Q table (Q2, Q6)
// Row = [key+1 as u64 bits | A_max doubles]; open addressing, linear probing, no deletion.  A row is created
// exactly where the reference's __check_entry (distr_q.py:47-57) would insert a dict entry, so the exported
// key set equals the reference's.
SFL_FN int key_port(unsigned key) { return (int)(key / (unsigned)(c_L.NT * 48)); }

SFL_NI double *q_row(Env e, const sfl_hparams &hp, unsigned key) {
  unsigned mask = (unsigned)c_L.q_cap - 1u;
  unsigned i = (key * 2654435761u) >> 7;
  for (int probe = 0; probe < c_L.q_cap; probe++) {
    i &= mask;
    double *row = e.q() + (size_t)i * c_L.q_stride;
    unsigned long long k = *(unsigned long long *)row;
    if (k == (unsigned long long)key + 1ull) return row + 1;
    if (k == 0ull) {
      if (e.h()->q_rows >= c_L.q_cap - 1) break;
      *(unsigned long long *)row = (unsigned long long)key + 1ull;
      e.h()->q_rows++;
      int port = key_port(key);
      int A = c_m.sw[c_m.port_switch[port]].y;
      for (int a = 0; a < A; a++) row[1 + a] = hp.default_q;
      if (e.h()->q_init_on) {                                   // distr_q.py:81-181 (lazy: same values, created on first touch)
        unsigned rem = key % (unsigned)(c_L.NT * 48);
        int tgt = (int)(rem / 48u), semb = (int)((rem % 48u) / 3u);
        int qi = c_m.qinit[port * c_L.NT + tgt];
        if (qi >= 0 && semb != 0) row[1 + (qi & 15)] = (qi & 16) ? 1000.0 : 500.0;
      }
      return row + 1;
    }
    i++;
  }
  e.h()->err |= SFL_ERR_Q_FULL;
  return e.q() + 1;     // keep running on row 0 (flagged)
}

SFL_FN double decay_pow(double rate, int n) { return rate == 1.0 ? 1.0 : pow(rate, (double)n); }

// distr_q.py:419-447 update (fp64, Python operator order, no FMA contraction)
SFL_NI void q_update(Env e, const sfl_hparams &hp, unsigned key, int action, double reward,
                     int has_next, unsigned next_key, int prev_sw, int next_sw) {
  double *row = q_row(e, hp, key);
  double lr = dmul(hp.lr, decay_pow(hp.lr_decay_rate, e.ninter()[prev_sw]));
  double one_m = dadd(1.0, -lr);
  double q = row[action];
  if (next_sw != prev_sw) {
    double mq = 0.0;
    if (has_next) {                                             // distr_q.py:449-466 max_q ignores the mask
      double *nrow = q_row(e, hp, next_key);
      int A = c_m.sw[next_sw].y;
      mq = nrow[0];
      for (int a = 1; a < A; a++) mq = nrow[a] > mq ? nrow[a] : mq;
    }
    row[action] = dadd(dmul(one_m, q), dmul(lr, dadd(reward, dmul(hp.gamma, mq))));
  } else {
    row[action] = dadd(dmul(one_m, q), dmul(lr, reward));
  }
}

// distr_q.py:468-490 max_action
SFL_FN int max_action(const double *row, int A, int mask) {
  int best = 0;
  for (int a = 1; a < A; a++) if (row[a] > row[best]) best = a;
  if ((mask >> best) & 1) return best;
  int b2 = -1;
  for (int a = 0; a < A; a++) if (((mask >> a) & 1) && (b2 < 0 || row[a] > row[b2])) b2 = a;
  return b2;
}

// ------------------------------------------------------------------------------------------------ E3
SFL_FN void sem_delete_owned(Env e, int port, int h) {
  int4 sw = c_m.sw[c_m.port_switch[port]];
  for (int k = 0; k < sw.x; k++) {
    int p = sw.z + k;
    if (e.sem()[p].z == h) e.sem()[p].z = -1;
  }
}

// rail_network.py:303-416 transition_semaphore, step by step
SFL_NI void transition_semaphore(Env e, int source, int out_port, int target, int h, int now) {
  if (e.state()[h] != ST_MALF) {                                          // :315-323
    sem_delete_owned(e, e.next_port()[h], h);
    if (e.prev_port()[h] >= 0) sem_delete_owned(e, e.prev_port()[h], h);
  }
  int4 r = e.sem()[out_port];                                             // :326-334
  if (r.z < 0) e.sem()[out_port] = make_int4(now, now + 3, h, SEM_OUT);
  else if (r.w == SEM_OUT || r.x > now) e.sem()[out_port] = make_int4(now, now + 3, h, r.w);
  int d_ot = c_m.port[out_port].y;
  r = e.sem()[target];                                                    // :336-344
  if (r.z < 0) e.sem()[target] = make_int4(now, now + d_ot + 1, h, SEM_IN);
  else if (r.w == SEM_IN || r.x > now) e.sem()[target] = make_int4(now, now + d_ot + 1, h, r.w);
  int4 tp = c_m.port[target];
  if (tp.z == 1) {                                                      // :356 forced path through the next switch
    int unique = tp.w;
    int4 up = c_m.port[unique];
    int far_port = up.x;
    if (unique != source && unique != out_port && unique != target) {   // :368-378
      r = e.sem()[unique];
      if (r.z < 0 || r.w == SEM_OUT || r.x > now) e.sem()[unique] = make_int4(now, now + d_ot + 1, h, SEM_OUT);
    }
    r = e.sem()[unique];                                                  // :380-388 (the list == 'out' test is never true)
    if (r.z < 0 || r.x > now) e.sem()[unique] = make_int4(now, now + d_ot, h, SEM_OUT);
    if (far_port != source && far_port != out_port && far_port != unique) {   // :390-402
      r = e.sem()[far_port];
      if (r.z < 0 || r.w == SEM_IN || r.x > now) e.sem()[far_port] = make_int4(now, now + d_ot + up.y + 1, h, SEM_IN);
    }
  }
  if (target != source && target != out_port) {                         // :404-414 moving edge
    r = e.sem()[target];
    if (r.z < 0 || r.w == SEM_OUT || r.x > now) e.sem()[target] = make_int4(now, now + d_ot + 1, h, SEM_OUT);
  }
}

// ------------------------------------------------------------------------------------------------ decision (lane 0)
SFL_FN int delay_at(Env e, int t, int cell, int dir, int now, int la) {
  int d = c_m.dist[((size_t)c_m.train0[t].w * (c_m.Hp * c_m.Wp) + cell) * 4 + dir];
  if (d >= SFL_INF_DIST) { e.h()->err |= SFL_ERR_INF_DISTANCE; d = 0; }       // observer.py:35-36
  return now - la + d;                                                      // observer.py:41
}

SFL_FN void pend_put(Env e, int t, int next_sw, unsigned key, int action, int prev_sw) {
  // distr_q.py:340-342 update_dict[(next_switch, train)] = (obs, action, agent): same key overwrites in place
  int n = e.pend_n()[t];
  unsigned meta = (unsigned)next_sw | ((unsigned)prev_sw << 12) | ((unsigned)action << 24);
  for (int i = 0; i < n; i++)
    if ((e.pend_meta()[t * c_L.pend_cap + i] & 0xFFFu) == (unsigned)next_sw) {
      e.pend_key()[t * c_L.pend_cap + i] = key; e.pend_meta()[t * c_L.pend_cap + i] = meta; return;
    }
  if (n >= c_L.pend_cap) { e.h()->err |= SFL_ERR_PEND_FULL; return; }
  e.pend_key()[t * c_L.pend_cap + n] = key; e.pend_meta()[t * c_L.pend_cap + n] = meta; e.pend_n()[t] = (uint8_t)(n + 1);
}

// the tail of one iteration of distr_q.py:302-362 that must wait for the train ticks run inside env.step()
// (switch_env.py:648-649): arrival flush (:345-356), interaction counter (:362), truncation (switch_env.py:652-657)
SFL_NI void finish_decision(Env e, const sfl_hparams &hp, int env_id) {
  EnvHdr *h = e.h();
  int learning = c_ra.mode != SFL_MODE_GREEDY;
  if (learning) {
    unsigned long long fresh = h->done_mask & ~h->at_dest_mask;
    while (fresh) {
      int t = ffs64(fresh); fresh &= fresh - 1;
      h->at_dest_mask |= 1ull << t;
      for (int i = 0; i < e.pend_n()[t]; i++) {
        unsigned meta = e.pend_meta()[t * c_L.pend_cap + i];
        q_update(e, hp, e.pend_key()[t * c_L.pend_cap + i], (int)(meta >> 24) & 15, 1000.0, 0, 0u, (int)(meta >> 12) & 0xFFF, -1);
      }
      e.pend_n()[t] = 0;
    }
    e.ninter()[h->pending_fin]++;
  }
  h->pending_fin = -1;
  if (h->step_counter > c_ra.max_steps) h->truncated = 1;
  if (c_ra.trace_dec && h->cur_dec >= 0 && h->cur_dec < c_ra.dec_cap) {
    sfl_dec_rec *rec = c_ra.trace_dec + (size_t)env_id * c_ra.dec_cap + h->cur_dec;
    rec->arrived = h->done_mask;
    rec->done = h->terminated | (h->truncated << 1);
    if (c_ra.trace_sem_buf) {
      int4 *dst = c_ra.trace_sem_buf + ((size_t)env_id * c_ra.dec_cap + h->cur_dec) * c_L.NP;
      for (int p = 0; p < c_L.NP; p++) dst[p] = e.sem()[p];
    }
  }
  h->cur_dec = -1;
}

// one switch-agent decision: observe (O1-O3) -> act (Q1) -> apply (E2, E3, R1) -> Q-update (Q2, Q3)
SFL_NI void decide(Env e, const sfl_hparams &hp, int env_id, int t) {
  EnvHdr *h = e.h();
  const int now = h->elapsed;
  const int s = e.act_switch()[t];
  const int4 sw = c_m.sw[s];
  const int P = sw.x, A = sw.y, p0 = sw.z, a0 = sw.w;
  const int4 tr0 = c_m.train0[t], tr1 = c_m.train1[t];
  // ---- observe (observer.py:246-308)
  int semb = 0, cur = -1;
  const int my_port = e.next_port()[t];
  for (int k = 0; k < P; k++) {
    int port = p0 + k;
    if (!port_blocked(e, c_m.port[port].x, port, t, now)) semb |= 1 << k;
    if (my_port == port) cur = k;
  }
  if (cur < 0) { h->err |= SFL_ERR_NO_TRAIN_AT_SWITCH; cur = 0; }
  int delay = delay_at(e, t, e.pos()[t], e.dir()[t], now, tr1.y);
  int level = delay <= 0 ? 0 : (delay <= (tr1.y - tr1.x) * 20 ? 1 : 2);           // observer.py:239-244
  unsigned key = (((unsigned)(p0 + cur) * c_L.NT + tr0.w) * 16u + semb) * 3u + level;
  int mask = 1 << (A - 1);                                                        // switch_agents.py:104-134
  for (int a = 0; a < A - 1; a++) {
    int4 ac = c_m.act[a0 + a];
    if (ac.x == cur && ((semb >> ac.y) & 1)) mask |= 1 << a;
  }
  const int reward_in = e.rewards()[s * c_L.T + t];                                   // last(): _cumulative_rewards[agent][train]
  // ---- act (distr_q.py:312-320 / :211)
  int action;
  if (c_ra.mode == SFL_MODE_REPLAY) {
    if (h->act_cursor >= c_ra.act_cap) { h->err |= SFL_ERR_REPLAY_UNDERRUN; action = A - 1; }
    else action = c_ra.replay_act[(size_t)env_id * c_ra.act_cap + h->act_cursor++];
    if (action < 0 || action >= A) { h->err |= SFL_ERR_BAD_ACTION; action = A - 1; }
  } else if (c_ra.mode == SFL_MODE_GREEDY) {
    action = max_action(q_row(e, hp, key), A, mask);
  } else {
    double eps = dmul(hp.epsilon, decay_pow(hp.epsilon_decay_rate, e.ninter()[s]));
    U4 u = philox4x32((unsigned)h->step_counter, (unsigned)(hp.episode_base + h->episode), 0x5F1u, 0u, (unsigned)hp.seed, (unsigned)(hp.seed >> 32));
    double u01 = ((double)u.x + 0.5) * (1.0 / 4294967296.0);
    if (u01 < eps) {
      int nvalid = popc32((unsigned)mask);
      int pick = (int)(((unsigned long long)u.y * (unsigned)nvalid) >> 32);
      action = 0;
      for (int a = 0; a < A; a++) if ((mask >> a) & 1) { if (pick == 0) { action = a; break; } pick--; }
    } else {
      action = max_action(q_row(e, hp, key), A, mask);
    }
  }
  // ---- apply (switch_env.py:203-294, switch_agents.py:136-168)
  int moving = 0, move2 = A_STOP, in_port = my_port, out_port = my_port;
  if (action != A - 1) {
    int4 ac = c_m.act[a0 + action];
    if (ac.x == cur) { moving = 1; move2 = ac.z; in_port = p0 + ac.x; out_port = p0 + ac.y; }
  }
  int next_switch = s, next_port = -1;
  if (moving) {                                                                   // rail_network.py:246-278
    next_port = c_m.port[out_port].x;
    transition_semaphore(e, in_port, out_port, next_port, t, now);
    e.source_port()[t] = (int16_t)in_port;
    e.next_port()[t] = (int16_t)next_port;
    e.prev_port()[t] = (int16_t)out_port;
    next_switch = c_m.port_switch[next_port];
  }
  uint8_t *plan = e.plan() + t * SFL_PLAN_CAP;
  int pl = e.plan_len()[t];
  if (moving && pl > 0) { pl = 1; plan[1] = (uint8_t)move2; pl = 2; }             // :257-266
  else if (!moving) {                                                             // :267-270
    if (pl >= SFL_PLAN_CAP) { h->err |= SFL_ERR_PLAN_FULL; pl = SFL_PLAN_CAP - 1; }
    for (int i = pl; i > 0; i--) plan[i] = plan[i - 1];
    plan[0] = A_STOP; pl++;
  } else { plan[0] = A_FWD; plan[1] = (uint8_t)move2; pl = 2; }                   // :271-272
  e.plan_len()[t] = (uint8_t)pl;
  int all_blocked = 1;                                                            // :274-282
  if (moving) all_blocked = port_blocked(e, next_port, out_port, t, now);
  else
    for (int a = 0; a < A - 1; a++) {
      int4 ac = c_m.act[a0 + a];
      if (p0 + ac.x == in_port && !port_blocked(e, c_m.port[p0 + ac.y].x, p0 + ac.y, t, now)) all_blocked = 0;
    }
  int cell = e.pos()[t], dir = e.dir()[t];                                            // reward_func.py:23-78
  for (int i = 0; i < pl; i++)
    if (plan[i] != A_STOP) { Chk c = check_action(plan[i], cell, dir); cell = c.cell; dir = c.dir; }
  int curr = delay_at(e, t, cell, dir, now, tr1.y);
  int reward_out = e.last_delay()[t] - curr;
  if (!all_blocked && plan[0] == A_STOP) reward_out -= 1300;
  e.rewards()[next_switch * c_L.T + t] = reward_out;                                  // switch_env.py:289
  e.last_delay()[t] = curr;                                                         // switch_env.py:291
  h->step_counter++;
  // ---- learn (distr_q.py:329-342)
  if (c_ra.mode != SFL_MODE_GREEDY) {
    int n = e.pend_n()[t];
    for (int i = 0; i < n; i++) {
      unsigned meta = e.pend_meta()[t * c_L.pend_cap + i];
      if ((int)(meta & 0xFFFu) == s) {
        q_update(e, hp, e.pend_key()[t * c_L.pend_cap + i], (int)(meta >> 24) & 15, (double)reward_in, 1, key, (int)(meta >> 12) & 0xFFF, s);
        for (int j = i; j + 1 < n; j++) {
          e.pend_key()[t * c_L.pend_cap + j] = e.pend_key()[t * c_L.pend_cap + j + 1];
          e.pend_meta()[t * c_L.pend_cap + j] = e.pend_meta()[t * c_L.pend_cap + j + 1];
        }
        e.pend_n()[t] = (uint8_t)(n - 1);
        break;
      }
    }
    pend_put(e, t, next_switch, key, action, s);
  }
  h->cum_reward += (double)reward_in;                                             // distr_q.py:360
  if (c_ra.trace_dec) {
    h->cur_dec = h->n_dec_logged;
    if (h->n_dec_logged < c_ra.dec_cap) {
      sfl_dec_rec *rec = c_ra.trace_dec + (size_t)env_id * c_ra.dec_cap + h->n_dec_logged;
      rec->ep = h->episode; rec->tick = now; rec->sw = s; rec->train = t; rec->key = key; rec->mask = mask;
      rec->action = action; rec->next_sw = next_switch; rec->reward = reward_in; rec->done = 0; rec->arrived = 0;
    }
    h->n_dec_logged++;
  }
  h->decisions++;
  h->pending_fin = s;
}

// ------------------------------------------------------------------------------------------------ reset (E1)
// switch_env.py:93-158 with a constant map: state re-init + the precomputed _init_ports table (:507-568)
SFL_NI void env_reset(Env e, int lane) {
  EnvHdr *h = e.h();
  for (int t = lane; t < c_L.T; t += SFL_LANES) {
    e.pos()[t] = -1; e.dir()[t] = (uint8_t)c_m.train0[t].y; e.state()[t] = ST_WAITING; e.saved()[t] = 0; e.prev_act()[t] = A_NONE;
    e.plan_len()[t] = 0; e.malf()[t] = 0; e.next_port()[t] = (int16_t)c_m.train1[t].z; e.act_switch()[t] = -1; e.pend_n()[t] = 0;
    e.last_delay()[t] = c_m.init_delay[t];
    // prev_port / source_port are NOT cleared: RailNetwork.reset (rail_network.py:135-149) keeps them
  }
  for (int p = lane; p < c_L.NP; p += SFL_LANES) e.sem()[p] = make_int4(0, 0, -1, 0);
  for (int i = lane; i < c_L.S * c_L.T; i += SFL_LANES) e.rewards()[i] = 0;
  w_sync();
  if (lane == 0) {
    for (int t = 0; t < c_L.T; t++) {                                     // switch_env.py:564-568, train order
      int4 tr1 = c_m.train1[t];
      e.sem()[tr1.z] = make_int4(tr1.x - 2, tr1.x + tr1.w, t, SEM_IN);
    }
    h->elapsed = 0; h->step_counter = 0; h->num_malf = 0; h->terminated = 0; h->truncated = 0; h->need_reset = 0;
    h->pending_fin = -1; h->cur_dec = -1; h->ev_cursor = 0; h->active_mask = 0; h->malf_prev_mask = 0; h->at_dest_mask = 0;
    h->done_mask = 0; h->cum_reward = 0.0;
  }
  w_sync();
}

// ------------------------------------------------------------------------------------------------ one tick (E5-E7, F2-F5)
SFL_NI void env_tick(Env e, Scratch &sc, const sfl_hparams &hp, int env_id, int lane) {
  EnvHdr *h = e.h();
  const int T = c_L.T;
  const int now = h->elapsed + 1;                                        // flatland: _elapsed_steps += 1 first
  const int replay_ev = c_ra.replay_ev != 0;
  if (replay_ev) {
    for (int t = lane; t < T; t += SFL_LANES) sc.inj[t] = 0;
    w_sync();
    if (lane == 0) {
      const int *ev = c_ra.replay_ev + (size_t)env_id * c_ra.ev_cap * 3;
      int c = h->ev_cursor;
      while (c < c_ra.ev_cap && ev[c * 3] >= 0 && ev[c * 3] <= now) { if (ev[c * 3] == now) sc.inj[ev[c * 3 + 1]] = ev[c * 3 + 2]; c++; }
      h->ev_cursor = c;
    }
    w_sync();
  }
  // ---- phase A: per train: plan pop (switch_env.py:304-339) + flatland step part 1 (Appendix B step 2)
  for (int t = lane; t < T; t += SFL_LANES) {
    int st = e.state()[t], p = e.pos()[t], d = e.dir()[t];
    int a = A_NOTHING, flags = 0, ecell = -1;
    if (st != ST_DONE) {
      int pl = e.plan_len()[t];
      uint8_t *plan = e.plan() + t * SFL_PLAN_CAP;
      if (pl == 0) a = A_FWD;
      else { a = plan[0]; e.prev_act()[t] = (uint8_t)a; for (int i = 1; i < pl; i++) plan[i - 1] = plan[i]; e.plan_len()[t] = (uint8_t)(pl - 1); }
      if (p >= 0) { Chk c = check_action(a, p, d); ecell = c.valid ? c.cell : p; flags = 1 | (c.valid ? 2 : 0); }
    }
    sc.act[t] = (uint8_t)a; sc.exp_cell[t] = ecell; sc.exp_flags[t] = (uint8_t)flags;
    // F5 malfunction draw: every train, every tick; applied only when the counter is 0
    int dur;
    if (replay_ev) dur = sc.inj[t];
    else {
      dur = 0;
      if (hp.malf_threshold) {
        U4 u = philox4x32((unsigned)now, (unsigned)t, 0xA11Fu, 0u, (unsigned)hp.seed, (unsigned)(hp.seed >> 32));
        if (u.x < hp.malf_threshold) dur = hp.malf_min + (int)(((unsigned long long)u.y * (unsigned)(hp.malf_max - hp.malf_min + 1)) >> 32) + 1;
      }
    }
    int mc = e.malf()[t];
    if (mc == 0 && dur > 0) { mc = dur; e.malf()[t] = (int16_t)mc; }
    // action preprocessing
    int act = a, saved = e.saved()[t];
    if (act == A_NOTHING) act = (st == ST_MOVING) ? A_FWD : (saved ? saved : A_STOP);
    if (st == ST_WAITING) act = A_NOTHING;
    int4 tr0 = c_m.train0[t];
    int pp = p >= 0 ? p : tr0.x, dd = p >= 0 ? d : tr0.y;
    if (act == A_LEFT || act == A_RIGHT) { Chk c = check_action(act, pp, dd); if (!(c.cell_ok && c.valid)) act = A_FWD; }
    if (is_moving(act)) { Chk c = check_action(act, pp, dd); if (!(c.cell_ok && c.valid)) act = A_STOP; }
    if (is_moving(act) && !saved && st != ST_DONE) { saved = act; e.saved()[t] = (uint8_t)saved; }
    int upd = (mc == 0) && act != A_STOP;
    int ncell = p, nd = d;
    if (st == ST_DONE) { }
    else if (p < 0 && saved) { ncell = tr0.x; nd = tr0.y; }
    else if (saved && upd) { Chk c = check_action(saved, p, d); if (c.cell_ok && c.valid) { ncell = c.cell; nd = c.dir; } act = saved; }
    int src = p >= 0 ? p : -1 - t;
    sc.src[t] = src; sc.dst[t] = ncell >= 0 ? ncell : src; sc.ndir[t] = nd; sc.pre[t] = (uint8_t)act;
  }
  w_sync();
  // ---- phase B: motion check (F3)
  for (int t = lane; t < T; t += SFL_LANES) {
    int s = sc.src[t], d = sc.dst[t];
    int wants = d != s, blocked = !wants, occ = -1;
    if (wants) {
      for (int k = 0; k < T; k++) {
        if (k == t) continue;
        int sk = sc.src[k], dk = sc.dst[k];
        if (sk == d) occ = k;
        if (dk == d && dk != sk && k < t) blocked = 1;                 // lowest handle wins a contended cell
      }
      if (occ >= 0 && sc.dst[occ] == s && sc.dst[occ] != sc.src[occ]) blocked = 1;   // swap
    }
    sc.occ[t] = occ; sc.blk[t] = (uint8_t)blocked;
  }
  w_sync();
  for (int iter = 0; iter < T; iter++) {                                // chains: fixed point (monotone)
    int changed = 0;
    for (int t = lane; t < T; t += SFL_LANES) {
      int occ = sc.occ[t];
      if (!sc.blk[t] && occ >= 0 && ((volatile uint8_t *)sc.blk)[occ]) { sc.blk[t] = 1; changed = 1; }
    }
    w_sync();
    if (!w_any(changed)) break;
  }
  // ---- phase C: state machine + position update (Appendix B steps 4-5)
  unsigned long long done_bits = 0, malf_bits = 0, stopped_bits = 0, depart_bits = 0;
  int on_map = 0;
  for (int t = lane; t < T; t += SFL_LANES) {
    int st = e.state()[t], mc = e.malf()[t], act = sc.pre[t];
    int wants = sc.dst[t] != sc.src[t];
    int in_malf = mc > 0, allowed = !in_malf && wants && !sc.blk[t];
    int4 tr0 = c_m.train0[t], tr1 = c_m.train1[t];
    int ed_reached = now >= tr1.x, stop_given = act == A_STOP, valid_move = is_moving(act) && allowed, conflict = !allowed;
    int nxt = st;
    switch (st) {
      case ST_WAITING: if (in_malf) nxt = ST_MALF_OFF; else if (ed_reached) nxt = ST_READY; break;
      case ST_READY: if (in_malf) nxt = ST_MALF_OFF; else if (valid_move) nxt = ST_MOVING; break;
      case ST_MALF_OFF:
        if (!in_malf) { if (ed_reached) nxt = valid_move ? ST_MOVING : (stop_given ? ST_STOPPED : ST_READY); else nxt = ST_WAITING; }
        break;
      case ST_MOVING: if (in_malf) nxt = ST_MALF; else if (stop_given || conflict) nxt = ST_STOPPED; break;
      case ST_STOPPED: if (in_malf) nxt = ST_MALF; else if (valid_move) nxt = ST_MOVING; break;
      case ST_MALF: if (!in_malf && valid_move) nxt = ST_MOVING; else if (!in_malf && (stop_given || conflict)) nxt = ST_STOPPED; break;
      default: break;
    }
    int p = e.pos()[t], d = e.dir()[t];
    if (nxt >= ST_MOVING && nxt <= ST_MALF) {
      if (st <= ST_MALF_OFF) { p = tr0.x; d = tr0.y; }
      else if (allowed) { p = sc.dst[t]; d = sc.ndir[t]; if (p == tr0.z) nxt = ST_DONE; }
    }
    if (nxt == ST_DONE) p = -1;
    if (mc > 0) mc--;
    e.state()[t] = (uint8_t)nxt; e.pos()[t] = p; e.dir()[t] = (uint8_t)d; e.malf()[t] = (int16_t)mc;
    if (p >= 0) e.saved()[t] = 0;
    if (nxt == ST_DONE) done_bits |= 1ull << t; else on_map++;
    if (mc != 0) malf_bits |= 1ull << t;
    if (nxt == ST_STOPPED || nxt == ST_MALF) stopped_bits |= 1ull << t;
    if (now == tr1.x - 2) depart_bits |= 1ull << t;
    if (c_ra.trace_tick && h->n_tick_logged < c_ra.tick_cap) {
      sfl_tick_rec *rec = c_ra.trace_tick + ((size_t)env_id * c_ra.tick_cap + h->n_tick_logged) * T + t;
      rec->pos = p; rec->dir = (int8_t)d; rec->state = (int8_t)nxt; rec->malf = (int16_t)mc;
    }
    // ---- phase D1: flatland held the train back (switch_env.py:353-367)
    int fl = sc.exp_flags[t];
    if ((fl & 1) && (fl & 2) && sc.exp_cell[t] != p && sc.act[t] != A_STOP) {
      int pl = e.plan_len()[t];
      uint8_t *plan = e.plan() + t * SFL_PLAN_CAP;
      if (pl >= SFL_PLAN_CAP) { h->err |= SFL_ERR_PLAN_FULL; pl = SFL_PLAN_CAP - 1; }
      for (int i = pl; i > 0; i--) plan[i] = plan[i - 1];
      plan[0] = sc.act[t]; e.plan_len()[t] = (uint8_t)(pl + 1);
      if (c_m.cell_switch[sc.exp_cell[t]] >= 0) e.next_port()[t] = e.source_port()[t];
    }
  }
  done_bits = w_or64(done_bits); malf_bits = w_or64(malf_bits); stopped_bits = w_or64(stopped_bits); depart_bits = w_or64(depart_bits);
  const unsigned long long all_mask = T >= 64 ? ~0ull : ((1ull << T) - 1ull);
  const int all_done = done_bits == all_mask;
  const int ended = all_done || now >= c_m.max_episode_steps;             // dones["__all__"] (Appendix B step 6)
  const unsigned long long prev_done = h->done_mask;
  w_sync();
  // ---- phase D2: semaphores of done trains (switch_env.py:370-376); every train counts as done at the end
  if ((done_bits & ~prev_done) || ended) {
    for (int p = lane; p < c_L.NP; p += SFL_LANES) {
      int tr = e.sem()[p].z;
      if (tr >= 0 && (ended || ((done_bits >> tr) & 1))) e.sem()[p].z = -1;
    }
    w_sync();
  }
  // ---- phase D3: departure bookings (switch_env.py:379-384), train order
  if (depart_bits && lane == 0) {
    unsigned long long b = depart_bits;
    while (b) {
      int t = ffs64(b); b &= b - 1;
      int4 tr1 = c_m.train1[t];
      e.sem()[e.next_port()[t]] = make_int4(tr1.x - 2, tr1.x + tr1.w, t, SEM_IN);
    }
  }
  w_sync();
  // ---- phase D4: extend_semaphores (rail_network.py:229-244)
  if (stopped_bits) {
    for (int p = lane; p < c_L.NP; p += SFL_LANES) {
      int4 r = e.sem()[p];
      if (r.z >= 0 && ((stopped_bits >> r.z) & 1)) { r.y = now + (r.y - r.x); r.x = now; e.sem()[p] = r; }
    }
    w_sync();
    if (lane == 0) {
      unsigned long long b = stopped_bits;
      while (b) {
        int t = ffs64(b); b &= b - 1;
        if (e.state()[t] == ST_MALF) {
          int port = e.next_port()[t];
          if (e.sem()[port].z < 0) e.sem()[port] = make_int4(now, now + c_m.train1[t].w, t, SEM_IN);
        }
      }
    }
    w_sync();
  }
  // ---- phase E: _check_active_switch (switch_env.py:427-485)
  unsigned long long active = 0;
  for (int t = lane; t < T; t += SFL_LANES) {
    int p = e.pos()[t], st = e.state()[t];
    if (p < 0 || st == ST_WAITING) continue;
    int peek = e.plan_len()[t] ? e.plan()[t * SFL_PLAN_CAP] : A_FWD;
    Chk c = check_action(peek, p, e.dir()[t]);
    int s = c_m.cell_switch[c.cell];
    if (s < 0) continue;
    if (st == ST_MOVING || st == ST_READY) { }
    else if ((st == ST_STOPPED || st == ST_MALF) && e.prev_act()[t] == A_STOP) { }
    else if (st == ST_STOPPED || st == ST_MALF) s = c_m.port_switch[e.next_port()[t]];
    else continue;
    e.act_switch()[t] = (int16_t)s;
    active |= 1ull << t;
  }
  active = w_or64(active);
  if (lane == 0) {
    h->elapsed = now;
    h->done_mask = done_bits;
    h->terminated = ended;
    h->num_malf += popc64(malf_bits & ~h->malf_prev_mask);             // switch_env.py:399-401
    h->malf_prev_mask = malf_bits;
    h->active_mask = active;
    h->ticks++;
    h->train_ticks += (unsigned long long)(T - popc64(prev_done));
    if (c_ra.trace_tick) h->n_tick_logged++;
  }
  w_sync();
}

// ------------------------------------------------------------------------------------------------ episode end
SFL_NI void episode_end(Env e, int env_id) {   // lane 0
  EnvHdr *h = e.h();
  if (c_ra.ep_log && h->n_ep_logged < c_ra.ep_cap) {
    sfl_ep_rec *rec = c_ra.ep_log + (size_t)env_id * c_ra.ep_cap + h->n_ep_logged;
    rec->cum_reward = h->cum_reward; rec->decisions = h->step_counter; rec->arrived = popc64(h->done_mask);
    rec->num_malfunctions = h->num_malf; rec->ticks = h->elapsed;
    if (c_ra.ep_delay) {
      int *d = c_ra.ep_delay + ((size_t)env_id * c_ra.ep_cap + h->n_ep_logged) * c_L.T;
      for (int t = 0; t < c_L.T; t++) d[t] = e.last_delay()[t];
    }
  }
  h->n_ep_logged++;
  h->episode++;
  h->need_reset = 1;
}

// ------------------------------------------------------------------------------------------------ the per-env driver
// `hot`/`hot_bytes`: per-warp staging area for the hot part of the env block (null on the host build);
// `hp_stage`: per-warp copy of the env's hyper-parameter record.
SFL_FN void env_run(Scratch &sc, int env_id, int lane,
                    char *hot, unsigned hot_bytes, sfl_hparams *hp_stage, int q_init_on) {
  char *gbase = c_ra.state + (size_t)env_id * c_L.env_stride;
  if (hot) {
    for (unsigned o = lane * 16u; o < hot_bytes; o += SFL_LANES * 16u) *(int4 *)(hot + o) = *(const int4 *)(gbase + o);
    for (unsigned o = lane * 4u; o < (unsigned)sizeof(sfl_hparams); o += SFL_LANES * 4u)
      *(int *)((char *)hp_stage + o) = *(const int *)((const char *)(c_ra.hp + env_id) + o);
    w_sync();
  }
  Env e;
  e.gb = gbase; e.hot = hot ? hot : gbase; e.semb = (hot && hot_bytes > c_L.off_sem) ? hot : gbase;
  EnvHdr *h = e.h();
  const sfl_hparams &hp = hot ? *hp_stage : c_ra.hp[env_id];
  if (lane == 0) h->q_init_on = q_init_on;
  w_sync();
  int budget = c_ra.max_ticks;
  for (;;) {
    if (h->halted) break;
    if (h->need_reset) {
      if (hp.episodes >= 0 && h->episode >= hp.episodes) { if (lane == 0) h->halted = 1; w_sync(); break; }
      env_reset(e, lane);
    }
    if (!h->terminated && !h->truncated && h->active_mask) {
      if (lane == 0) {
        while (h->active_mask && !h->truncated) {                       // agent_iter: FIFO in train-handle order
          int t = ffs64(h->active_mask);
          h->active_mask &= h->active_mask - 1;
          decide(e, hp, env_id, t);
          if (h->active_mask) finish_decision(e, hp, env_id); // no ticks follow this decision
        }
      }
      w_sync();
    }
    if (h->terminated || h->truncated) {
      if (lane == 0) episode_end(e, env_id);
      w_sync();
      continue;
    }
    if (budget == 0) break;
    env_tick(e, sc, hp, env_id, lane);
    budget--;
    if (h->active_mask || h->terminated) {
      if (lane == 0 && h->pending_fin >= 0) finish_decision(e, hp, env_id);
      w_sync();
    }
  }
  if (lane == 0) {
    sfl_env_counters *c = c_ra.counters + env_id;
    c->decisions = h->decisions; c->ticks = h->ticks; c->train_ticks = h->train_ticks; c->episodes = h->episode;
    c->err = h->err; c->q_rows = h->q_rows; c->halted = h->halted; c->n_dec_logged = h->n_dec_logged;
    c->n_tick_logged = h->n_tick_logged; c->n_ep_logged = h->n_ep_logged; c->elapsed = h->elapsed;
  }
  if (hot) {
    w_sync();
    for (unsigned o = lane * 16u; o < hot_bytes; o += SFL_LANES * 16u) *(int4 *)(gbase + o) = *(const int4 *)(hot + o);
  }
}

}  // namespace sfl
