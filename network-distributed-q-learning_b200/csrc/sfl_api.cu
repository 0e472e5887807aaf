// sfl_api.cu -- kernels + C-ABI (include/switchfl_b200.h) of the B200 SwitchFL backend.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a (see __graft_entry__.build).  There is no
// CPU path in that build.  tests/emul compiles this same file with g++ -DSFL_HOST_EMUL to unit-test the
// per-environment logic where no GPU exists; that library is never loaded by the product package.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "sfl_core.cuh"

using namespace sfl;

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, const char *detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

// ------------------------------------------------------------------------------------------------ backend glue
#ifndef SFL_HOST_EMUL
#include <cuda_runtime.h>
#define CU(call)                                                               \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) return fail(SFL_E_CUDA, "CUDA: %s", cudaGetErrorString(e_)); \
  } while (0)
#define CK(call)                                                               \
  do {                                                                         \
    if (call) return fail(SFL_E_CUDA, "CUDA: %s", cudaGetErrorString(cudaGetLastError())); \
  } while (0)
static int dev_alloc(void **p, size_t n) { return cudaMalloc(p, n) == cudaSuccess ? 0 : 1; }
static void dev_free(void *p) { cudaFree(p); }
static int h2d(void *d, const void *h, size_t n, void *s) { return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)s) != cudaSuccess; }
static int d2h(void *h, const void *d, size_t n, void *s) {
  if (cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)s) != cudaSuccess) return 1;
  return cudaStreamSynchronize((cudaStream_t)s) != cudaSuccess;
}
static int dev_zero(void *d, size_t n, void *s) { return cudaMemsetAsync(d, 0, n, (cudaStream_t)s) != cudaSuccess; }
#else
#define CU(call) do { if (call) return fail(SFL_E_CUDA, "emul: %s", #call); } while (0)
#define CK(call) CU(call)
static int dev_alloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p == nullptr; }
static void dev_free(void *p) { free(p); }
static int h2d(void *d, const void *h, size_t n, void *) { memcpy(d, h, n); return 0; }
static int d2h(void *h, const void *d, size_t n, void *) { memcpy(h, d, n); return 0; }
static int dev_zero(void *d, size_t n, void *) { memset(d, 0, n); return 0; }
#endif

#define SFL_WARPS_PER_CTA 4

// ------------------------------------------------------------------------------------------------ kernels
struct InitArgs { char *state; int n_envs, keep_q, keep_ninter, pad; };

SFL_FN void env_init(const InitArgs &ia, int env_id, int lane) {
  Env e;
  e.gb = e.hot = e.semb = ia.state + (size_t)env_id * c_L.env_stride;
  for (int t = lane; t < c_L.T; t += SFL_LANES) { e.prev_port()[t] = -1; e.source_port()[t] = -1; e.pend_n()[t] = 0; }
  if (!ia.keep_ninter) for (int s = lane; s < c_L.S; s += SFL_LANES) e.ninter()[s] = 0;
  if (!ia.keep_q) {
    size_t n = (size_t)c_L.q_cap * c_L.q_stride;
    for (size_t i = lane; i < n; i += SFL_LANES) e.q()[i] = 0.0;
  }
  if (lane == 0) {
    int q_rows = ia.keep_q ? e.h()->q_rows : 0;
    EnvHdr z;
    memset(&z, 0, sizeof(z));
    z.need_reset = 1; z.pending_fin = -1; z.cur_dec = -1; z.q_rows = q_rows;
    *e.h() = z;
  }
}

#ifndef SFL_HOST_EMUL
__global__ void __launch_bounds__(32 * SFL_WARPS_PER_CTA) k_init(InitArgs ia) {
  int env_id = blockIdx.x * SFL_WARPS_PER_CTA + (threadIdx.x >> 5);
  if (env_id < ia.n_envs) env_init(ia, env_id, threadIdx.x & 31);
}

// One warp per environment, SFL_WARPS_PER_CTA environments per CTA; the per-warp exchange block lives in
// shared memory.  Each warp advances its environment by up to max_ticks flatland ticks and every
// switch-agent decision in between, resetting the environment in place when an episode ends.
// Dynamic shared memory per warp: [hot env state (hot_bytes) | sfl_hparams | Scratch]; the hot state (header, train
// arrays, pending lists, semaphores when they fit) is staged once per launch and written back at the end, so the
// tick / decision loops touch HBM only for Q rows, the reward matrix and the interaction counters.
__global__ void __launch_bounds__(32 * SFL_WARPS_PER_CTA, 7) k_run(int q_init_on, unsigned hot_bytes, unsigned warp_smem) {
  extern __shared__ __align__(16) char smem[];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int env_id = blockIdx.x * SFL_WARPS_PER_CTA + w;
  if (env_id >= c_ra.n_envs) return;
  char *mine = smem + (size_t)w * warp_smem;
  sfl_hparams *hp_stage = (sfl_hparams *)(mine + hot_bytes);
  Scratch *sc = (Scratch *)(mine + hot_bytes + ((sizeof(sfl_hparams) + 15) / 16 * 16));
  env_run(*sc, env_id, lane, mine, hot_bytes, hp_stage, q_init_on);
}

__global__ void k_sum(const sfl_env_counters *c, int n, unsigned long long *out) {
  unsigned long long d = 0, t = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { d += c[i].decisions; t += c[i].ticks; }
  for (int o = 16; o; o >>= 1) { d += __shfl_down_sync(0xffffffffu, d, o); t += __shfl_down_sync(0xffffffffu, t, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, d); atomicAdd(out + 1, t); }
}
#endif

// ------------------------------------------------------------------------------------------------ context
struct Ctx {
  DevMap m;
  Layout L;
  sfl_config cfg;
  sfl_buffers bufs;
  int bound, device, q_init_on;
  unsigned hot_bytes, warp_smem;
  void *blob;          // device block holding every map table
  void *sum_buf;       // 2 x u64
  std::vector<int32_t> sw_A, port_switch;
};

// publish the context's map / layout (and the launch arguments) to constant memory, in stream order
struct Ctx;
static int set_constants(const Ctx *c, const RunArgs *ra, void *stream);

static unsigned align_up(unsigned v, unsigned a) { return (v + a - 1) / a * a; }

static int make_layout(const sfl_map_desc *map, const sfl_config *cfg, Layout *L, int *a_max_out) {
  if (!map || !cfg) return fail(SFL_E_ARG, "null argument%s");
  if (map->T < 1 || map->T > SFL_MAX_T) return fail(SFL_E_ARG, "T must be in 1..64%s");
  if (map->S < 1 || map->S > 4095) return fail(SFL_E_ARG, "S must be in 1..4095%s");
  if (map->NP > 32767) return fail(SFL_E_ARG, "too many ports%s");
  if (cfg->q_cap < 2 || (cfg->q_cap & (cfg->q_cap - 1))) return fail(SFL_E_ARG, "q_cap must be a power of two%s");
  if (cfg->pend_cap < 1 || cfg->pend_cap > 64 || cfg->n_envs < 1) return fail(SFL_E_ARG, "bad pend_cap / n_envs%s");
  int a_max = 0;
  for (int s = 0; s < map->S; s++) { if (map->sw_A[s] > a_max) a_max = map->sw_A[s]; if (map->sw_P[s] > 4) return fail(SFL_E_ARG, "P > 4%s"); }
  if (a_max > 15) return fail(SFL_E_ARG, "A > 15%s");
  if ((uint64_t)map->NP * map->NT * 48ull >= 0xFFFFFFFFull) return fail(SFL_E_ARG, "state index does not fit 32 bits%s");
  memset(L, 0, sizeof(*L));
  L->T = map->T; L->S = map->S; L->NP = map->NP; L->NT = map->NT; L->a_max = a_max; L->q_cap = cfg->q_cap;
  L->q_stride = 1 + a_max; L->pend_cap = cfg->pend_cap;
  unsigned o = align_up((unsigned)sizeof(EnvHdr), 16);
  unsigned T = (unsigned)map->T;
#define PUT(field, bytes) L->field = o; o = align_up(o + (unsigned)(bytes), 16)
  PUT(off_pos, 4 * T); PUT(off_last_delay, 4 * T);
  PUT(off_malf, 2 * T); PUT(off_next_port, 2 * T); PUT(off_prev_port, 2 * T); PUT(off_source_port, 2 * T); PUT(off_act_switch, 2 * T);
  PUT(off_dir, T); PUT(off_state, T); PUT(off_saved, T); PUT(off_prev_act, T); PUT(off_plan_len, T); PUT(off_plan, SFL_PLAN_CAP * T); PUT(off_pend_n, T);
  PUT(off_pend_key, 4 * T * cfg->pend_cap); PUT(off_pend_meta, 4 * T * cfg->pend_cap);
  PUT(off_sem, 16 * (unsigned)map->NP); PUT(off_rewards, 4 * (unsigned)map->S * T); PUT(off_ninter, 4 * (unsigned)map->S);
#undef PUT
  L->off_q = o;
  L->env_stride = ((unsigned long long)o + (unsigned long long)cfg->q_cap * L->q_stride * 8ull + 127ull) / 128ull * 128ull;
  *a_max_out = a_max;
  return SFL_OK;
}

static int set_constants(const Ctx *c, const RunArgs *ra, void *stream) {
#ifndef SFL_HOST_EMUL
  if (cudaMemcpyToSymbolAsync(c_m, &c->m, sizeof(DevMap), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbolAsync(c_L, &c->L, sizeof(Layout), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) return 1;
  if (ra && cudaMemcpyToSymbolAsync(c_ra, ra, sizeof(RunArgs), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) return 1;
#else
  c_m = c->m; c_L = c->L;
  if (ra) c_ra = *ra;
#endif
  return 0;
}

extern "C" {

int sfl_abi_version(void) { return SFL_ABI_VERSION; }
const char *sfl_last_error(void) { return g_err; }

int sfl_query_sizes(const sfl_map_desc *map, const sfl_config *cfg, sfl_sizes *out) {
  Layout L; int a_max;
  int rc = make_layout(map, cfg, &L, &a_max);
  if (rc) return rc;
  if (!out) return fail(SFL_E_ARG, "null out%s");
  uint64_t B = (uint64_t)cfg->n_envs;
  out->env_stride = L.env_stride;
  out->state_bytes = B * L.env_stride;
  out->hparams_bytes = B * sizeof(sfl_hparams);
  out->counters_bytes = B * sizeof(sfl_env_counters);
  out->trace_dec_bytes = B * (uint64_t)cfg->dec_cap * sizeof(sfl_dec_rec);
  out->trace_tick_bytes = B * (uint64_t)cfg->tick_cap * map->T * sizeof(sfl_tick_rec);
  out->trace_sem_bytes = cfg->trace_sem ? B * (uint64_t)cfg->dec_cap * map->NP * 16ull : 0;
  out->ep_log_bytes = B * (uint64_t)cfg->ep_cap * sizeof(sfl_ep_rec);
  out->ep_delay_bytes = B * (uint64_t)cfg->ep_cap * map->T * 4ull;
  out->replay_act_bytes = B * (uint64_t)cfg->act_cap;
  out->replay_ev_bytes = B * (uint64_t)cfg->ev_cap * 12ull;
  out->q_stride = L.q_stride;
  out->a_max = a_max;
  return SFL_OK;
}

int sfl_create(const sfl_map_desc *map, const sfl_config *cfg, int device, void **ctx_out) {
  if (!ctx_out) return fail(SFL_E_ARG, "null ctx%s");
  Layout L; int a_max;
  int rc = make_layout(map, cfg, &L, &a_max);
  if (rc) return rc;
#ifndef SFL_HOST_EMUL
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SFL_E_CUDA, "no CUDA device: this library has no CPU path%s");
  CU(cudaSetDevice(device));
#endif
  Ctx *c = new (std::nothrow) Ctx();
  if (!c) return fail(SFL_E_NOMEM, "host alloc%s");
  c->L = L; c->cfg = *cfg; c->bound = 0; c->device = device; c->q_init_on = 0; c->blob = nullptr; c->sum_buf = nullptr;
  memset(&c->bufs, 0, sizeof(c->bufs));
  const int H = map->H, W = map->W, Hp = H + 2, Wp = W + 2, T = map->T, NT = map->NT, S = map->S, NP = map->NP, NA = map->NA;
  auto pcell = [&](int cell) { return cell < 0 ? -1 : (cell / W + 1) * Wp + (cell % W + 1); };
  // ---- pack every table into one host image, 16-byte aligned sections
  std::vector<unsigned char> img;
  auto section = [&](size_t bytes) { size_t o = (img.size() + 15) / 16 * 16; img.resize(o + bytes, 0); return o; };
  size_t o_grid = section(2ull * Hp * Wp), o_csw = section(2ull * Hp * Wp), o_sw = section(16ull * S), o_port = section(16ull * NP);
  size_t o_psw = section(2ull * NP), o_act = section(16ull * (NA ? NA : 1)), o_t0 = section(16ull * T), o_t1 = section(16ull * T);
  size_t o_idl = section(4ull * T), o_dist = section(16ull * NT * Hp * Wp), o_qi = section((size_t)NP * NT);
  uint16_t *grid = (uint16_t *)&img[o_grid]; int16_t *csw = (int16_t *)&img[o_csw];
  for (int i = 0; i < Hp * Wp; i++) csw[i] = -1;
  for (int r = 0; r < H; r++) for (int cc = 0; cc < W; cc++) {
    grid[(r + 1) * Wp + cc + 1] = map->grid[r * W + cc];
    csw[(r + 1) * Wp + cc + 1] = (int16_t)map->cell_switch[r * W + cc];
  }
  int4 *sw = (int4 *)&img[o_sw];
  c->sw_A.assign(map->sw_A, map->sw_A + S);
  for (int s = 0; s < S; s++) sw[s] = make_int4(map->sw_P[s], map->sw_A[s], map->sw_port0[s], map->sw_act0[s]);
  int4 *port = (int4 *)&img[o_port]; int16_t *psw = (int16_t *)&img[o_psw];
  for (int p = 0; p < NP; p++) port[p] = make_int4(map->port_nbr[p], map->port_dist[p], map->port_n_intra[p], map->port_intra0[p]);
  for (int s = 0; s < S; s++) for (int p = map->sw_port0[s]; p < map->sw_port0[s + 1]; p++) psw[p] = (int16_t)s;
  c->port_switch.assign(NP, 0);
  for (int p = 0; p < NP; p++) c->port_switch[p] = psw[p];
  int4 *act = (int4 *)&img[o_act];
  for (int a = 0; a < NA; a++) act[a] = make_int4(map->act_in[a], map->act_out[a], map->act_move[a], 0);
  int4 *t0 = (int4 *)&img[o_t0], *t1 = (int4 *)&img[o_t1]; int *idl = (int *)&img[o_idl];
  for (int t = 0; t < T; t++) {
    t0[t] = make_int4(pcell(map->init_cell[t]), map->init_dir[t], pcell(map->target_cell[t]), map->tgt_index[t]);
    t1[t] = make_int4(map->ed[t], map->la[t], map->first_port[t], map->first_dist[t]);
    idl[t] = map->init_delay[t];
  }
  int *dist = (int *)&img[o_dist];
  for (size_t i = 0; i < (size_t)NT * Hp * Wp * 4; i++) dist[i] = SFL_INF_DIST;
  for (int k = 0; k < NT; k++) for (int r = 0; r < H; r++) for (int cc = 0; cc < W; cc++) for (int d = 0; d < 4; d++)
    dist[((size_t)k * Hp * Wp + (r + 1) * Wp + cc + 1) * 4 + d] = map->dist[(((size_t)k * H + r) * W + cc) * 4 + d];
  int8_t *qi = (int8_t *)&img[o_qi];
  for (int i = 0; i < NP * NT; i++) qi[i] = map->qinit_act[i] < 0 ? (int8_t)-1 : (int8_t)(map->qinit_act[i] | (map->qinit_final[i] ? 16 : 0));
  if (dev_alloc(&c->blob, img.size()) || dev_alloc(&c->sum_buf, 16)) { delete c; return fail(SFL_E_CUDA, "device alloc of map constants failed%s"); }
  if (h2d(c->blob, img.data(), img.size(), nullptr)) { delete c; return fail(SFL_E_CUDA, "upload of map constants failed%s"); }
#ifndef SFL_HOST_EMUL
  CU(cudaDeviceSynchronize());
#endif
  char *b = (char *)c->blob;
  DevMap &m = c->m;
  m.H = H; m.W = W; m.Hp = Hp; m.Wp = Wp; m.S = S; m.NP = NP; m.NA = NA; m.T = T; m.NT = NT; m.max_episode_steps = map->max_episode_steps;
  m.a_max = a_max; m.pad0 = 0;
  m.grid.p = (const uint16_t *)(b + o_grid); m.cell_switch.p = (const int16_t *)(b + o_csw); m.sw.p = (const int4 *)(b + o_sw);
  m.port.p = (const int4 *)(b + o_port); m.port_switch.p = (const int16_t *)(b + o_psw); m.act.p = (const int4 *)(b + o_act);
  m.train0.p = (const int4 *)(b + o_t0); m.train1.p = (const int4 *)(b + o_t1); m.init_delay.p = (const int *)(b + o_idl);
  m.dist.p = (const int *)(b + o_dist); m.qinit.p = (const int8_t *)(b + o_qi);
  // hot region: everything up to the reward matrix when the semaphore records fit the per-warp budget, else up to them
  c->hot_bytes = (L.off_rewards <= 8192u) ? L.off_rewards : L.off_sem;
  c->warp_smem = c->hot_bytes + (unsigned)((sizeof(sfl_hparams) + 15) / 16 * 16) + (unsigned)((sizeof(Scratch) + 15) / 16 * 16);
#ifndef SFL_HOST_EMUL
  CU(cudaFuncSetAttribute(k_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->warp_smem * SFL_WARPS_PER_CTA)));
#endif
  *ctx_out = c;
  return SFL_OK;
}

int sfl_destroy(void *ctx) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return SFL_OK;
  if (c->blob) dev_free(c->blob);
  if (c->sum_buf) dev_free(c->sum_buf);
  delete c;
  return SFL_OK;
}

int sfl_bind(void *ctx, const sfl_buffers *bufs) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !bufs) return fail(SFL_E_ARG, "null argument%s");
  if (!bufs->state || !bufs->hparams || !bufs->counters) return fail(SFL_E_ARG, "state, hparams and counters are mandatory%s");
  if (c->cfg.dec_cap > 0 && !bufs->trace_dec) return fail(SFL_E_ARG, "dec_cap > 0 needs trace_dec%s");
  if (c->cfg.tick_cap > 0 && !bufs->trace_tick) return fail(SFL_E_ARG, "tick_cap > 0 needs trace_tick%s");
  c->bufs = *bufs;
  c->bound = 1;
  return SFL_OK;
}

int sfl_reset(void *ctx, int keep, void *stream) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(SFL_E_ARG, "null ctx%s");
  if (!c->bound) return fail(SFL_E_STATE, "sfl_bind first%s");
  InitArgs ia; ia.state = (char *)c->bufs.state; ia.n_envs = c->cfg.n_envs; ia.keep_q = keep & 1; ia.keep_ninter = (keep >> 1) & 1; ia.pad = 0;
  CK(dev_zero(c->bufs.counters, (size_t)c->cfg.n_envs * sizeof(sfl_env_counters), stream));
#ifndef SFL_HOST_EMUL
  int grid = (c->cfg.n_envs + SFL_WARPS_PER_CTA - 1) / SFL_WARPS_PER_CTA;
  CK(set_constants(c, nullptr, stream));
  k_init<<<grid, 32 * SFL_WARPS_PER_CTA, 0, (cudaStream_t)stream>>>(ia);
  CU(cudaGetLastError());
#else
  set_constants(c, nullptr, stream);
  for (int i = 0; i < c->cfg.n_envs; i++) env_init(ia, i, 0);
#endif
  return SFL_OK;
}

int sfl_enable_q_init(void *ctx, int on) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(SFL_E_ARG, "null ctx%s");
  c->q_init_on = on ? 1 : 0;
  return SFL_OK;
}

int sfl_run(void *ctx, int mode, int max_ticks, void *stream) {
  Ctx *c = (Ctx *)ctx;
  if (!c) return fail(SFL_E_ARG, "null ctx%s");
  if (!c->bound) return fail(SFL_E_STATE, "sfl_bind first%s");
  if (mode < SFL_MODE_LEARN || mode > SFL_MODE_REPLAY || max_ticks < 0) return fail(SFL_E_ARG, "bad mode / max_ticks%s");
  if (mode == SFL_MODE_REPLAY && !c->bufs.replay_act) return fail(SFL_E_ARG, "replay mode needs replay_act%s");
  RunArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.mode = mode; ra.max_ticks = max_ticks; ra.n_envs = c->cfg.n_envs; ra.trace_sem = c->cfg.trace_sem;
  ra.dec_cap = c->cfg.dec_cap; ra.tick_cap = c->cfg.tick_cap; ra.ep_cap = c->cfg.ep_cap; ra.act_cap = c->cfg.act_cap;
  ra.ev_cap = c->cfg.ev_cap; ra.max_steps = c->cfg.max_steps;
  ra.state = (char *)c->bufs.state; ra.hp = (const sfl_hparams *)c->bufs.hparams; ra.counters = (sfl_env_counters *)c->bufs.counters;
  ra.trace_dec = c->cfg.dec_cap > 0 ? (sfl_dec_rec *)c->bufs.trace_dec : nullptr;
  ra.trace_tick = c->cfg.tick_cap > 0 ? (sfl_tick_rec *)c->bufs.trace_tick : nullptr;
  ra.trace_sem_buf = (c->cfg.trace_sem && c->cfg.dec_cap > 0) ? (int4 *)c->bufs.trace_sem : nullptr;
  ra.ep_log = c->cfg.ep_cap > 0 ? (sfl_ep_rec *)c->bufs.ep_log : nullptr;
  ra.ep_delay = c->cfg.ep_cap > 0 ? (int *)c->bufs.ep_delay : nullptr;
  ra.replay_act = (const int8_t *)c->bufs.replay_act;
  ra.replay_ev = (mode == SFL_MODE_REPLAY && c->cfg.ev_cap > 0) ? (const int *)c->bufs.replay_ev : nullptr;
#ifndef SFL_HOST_EMUL
  int grid = (c->cfg.n_envs + SFL_WARPS_PER_CTA - 1) / SFL_WARPS_PER_CTA;
  CK(set_constants(c, &ra, stream));
  k_run<<<grid, 32 * SFL_WARPS_PER_CTA, c->warp_smem * SFL_WARPS_PER_CTA, (cudaStream_t)stream>>>(c->q_init_on, c->hot_bytes, c->warp_smem);
  CU(cudaGetLastError());
#else
  static Scratch sc;
  set_constants(c, &ra, stream);
  for (int i = 0; i < c->cfg.n_envs; i++) env_run(sc, i, 0, nullptr, 0, nullptr, c->q_init_on);
#endif
  return SFL_OK;
}

int sfl_total_decisions(void *ctx, uint64_t *decisions, uint64_t *ticks, void *stream) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !c->bound) return fail(SFL_E_STATE, "sfl_bind first%s");
  unsigned long long out[2] = {0, 0};
#ifndef SFL_HOST_EMUL
  CK(dev_zero(c->sum_buf, 16, stream));
  k_sum<<<148, 256, 0, (cudaStream_t)stream>>>((const sfl_env_counters *)c->bufs.counters, c->cfg.n_envs, (unsigned long long *)c->sum_buf);
  CU(cudaGetLastError());
  CK(d2h(out, c->sum_buf, 16, stream));
#else
  const sfl_env_counters *cn = (const sfl_env_counters *)c->bufs.counters;
  for (int i = 0; i < c->cfg.n_envs; i++) { out[0] += cn[i].decisions; out[1] += cn[i].ticks; }
#endif
  if (decisions) *decisions = out[0];
  if (ticks) *ticks = out[1];
  return SFL_OK;
}

int sfl_export_q(void *ctx, int env, uint32_t *keys_host, double *vals_host, int cap_rows, int *n_rows, void *stream) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !c->bound) return fail(SFL_E_STATE, "sfl_bind first%s");
  if (env < 0 || env >= c->cfg.n_envs || !n_rows) return fail(SFL_E_ARG, "bad env / n_rows%s");
  const Layout &L = c->L;
  size_t n = (size_t)L.q_cap * L.q_stride;
  std::vector<double> img(n);
  CK(d2h(img.data(), (char *)c->bufs.state + (size_t)env * L.env_stride + L.off_q, n * 8, stream));
  int rows = 0;
  for (int i = 0; i < L.q_cap; i++) {
    unsigned long long k;
    memcpy(&k, &img[(size_t)i * L.q_stride], 8);
    if (!k) continue;
    if (rows < cap_rows && keys_host && vals_host) {
      keys_host[rows] = (uint32_t)(k - 1);
      for (int a = 0; a < L.a_max; a++) vals_host[(size_t)rows * L.a_max + a] = img[(size_t)i * L.q_stride + 1 + a];
    }
    rows++;
  }
  *n_rows = rows;
  return rows > cap_rows && keys_host ? fail(SFL_E_NOMEM, "export buffer too small%s") : SFL_OK;
}

int sfl_import_q(void *ctx, int env, const uint32_t *keys_host, const double *vals_host, int n_rows, void *stream) {
  Ctx *c = (Ctx *)ctx;
  if (!c || !c->bound) return fail(SFL_E_STATE, "sfl_bind first%s");
  if (env < 0 || env >= c->cfg.n_envs || n_rows < 0 || n_rows >= c->cfg.q_cap) return fail(SFL_E_ARG, "bad env / n_rows%s");
  const Layout &L = c->L;
  size_t n = (size_t)L.q_cap * L.q_stride;
  std::vector<double> img(n, 0.0);
  unsigned mask = (unsigned)L.q_cap - 1u;
  for (int r = 0; r < n_rows; r++) {
    unsigned key = keys_host[r];
    unsigned i = (key * 2654435761u) >> 7;
    for (;;) {
      i &= mask;
      unsigned long long k;
      memcpy(&k, &img[(size_t)i * L.q_stride], 8);
      if (k == 0 || k == (unsigned long long)key + 1ull) break;
      i++;
    }
    unsigned long long k = (unsigned long long)key + 1ull;
    memcpy(&img[(size_t)i * L.q_stride], &k, 8);
    for (int a = 0; a < L.a_max; a++) img[(size_t)i * L.q_stride + 1 + a] = vals_host[(size_t)r * L.a_max + a];
  }
  char *base = (char *)c->bufs.state + (size_t)env * L.env_stride;
  CK(h2d(base + L.off_q, img.data(), n * 8, stream));
  int q_rows = n_rows;
  CK(h2d(base + offsetof(EnvHdr, q_rows), &q_rows, 4, stream));
#ifndef SFL_HOST_EMUL
  CU(cudaStreamSynchronize((cudaStream_t)stream));
#endif
  return SFL_OK;
}

}  // extern "C"
