// pmaf_api.cu — C ABI (include/pmaf.h) over the sm_100a kernels. Host code only orchestrates:
// buffers, one stream per planner, launches, tiny H2D/D2H copies. There is no CPU compute path:
// every entry point that computes launches a kernel and fails with PMAF_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: libnccl is dlopen()ed when a planner is sharded

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstddef>
#include <cstring>
#include <ctime>
#include <random>
#include <string>
#include <vector>

#include "../../include/pmaf.h"
#include "pmaf_rollout.cuh"
#include "pmaf_dq.cuh"

using namespace pmaf;

// ---- error plumbing ------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(PMAF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define REQUIRE(cond, code, ...) \
  do {                           \
    if (!(cond)) return fail(code, __VA_ARGS__); \
  } while (0)

extern "C" const char *pmaf_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *pmaf_version(void) { return "pmaf 0.1 sm_100a fp64-exact -fmad=false"; }

// ---- planner object ----------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool owned = true;
  cudaError_t resize(size_t count) {
    if (count == n && p) return cudaSuccess;
    if (p && owned) cudaFree(p);
    p = nullptr, n = 0, owned = true;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void alias(T *ptr, size_t count) {  // a view into another allocation
    release();
    p = ptr, n = count, owned = false;
  }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr, n = 0, owned = true;
  }
};

struct pmaf_planner {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_roll[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [slot][begin/end]
  bool roll_timing[2] = {false, false};
  int roll_slot = 0;
  cudaEvent_t ev_d2h = nullptr, ev_stage = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  bool upload_dedup = true;
  bool time_rollouts = true;  // CUDA events around every rollout kernel (pmaf_set_rollout_timing)
  bool initialized = false;
  // shard
  int n_global = 0, first_agent = 0, rank = 0, world = 1;
  bool shard_set = false;
  void *nccl = nullptr;
  // configuration
  int A = 0, O = 0, H = 0;  // local agents, obstacles, max path points
  double goal[3] = {0, 0, 0}, mgr_init_pos[3] = {0, 0, 0};
  double delta_t = 0, pred_dt = 0, shell = 0, mass = 1, rad = 0, vel_max = 0, approach = 0;
  // device state
  DevBuf<double> k_attr, k_circ, k_repel, k_damp;  // [n_global]
  DevBuf<double> init_pos, cur_pos, vel, min_obs, path_len, ws_cost, pred_time, cost, paths, rot, random_vecs;
  DevBuf<int> n_path, reached;
  DevBuf<uint32_t> known;
  DevBuf<double> obs_pos, obs_vel, obs_rad;     // agents' obstacle copy at rollout start
  DevBuf<double> live_pos, live_vel, live_rad;  // list passed to moveRealEEAgent / resetEEAgents
  DevBuf<unsigned char> image, real_known;
  DevBuf<double> real_rot, best_random, scratch, real_path_out;
  DevBuf<RealState> real;
  DevBuf<DeviceBest> best;
  DevBuf<EvalResult> eval;
  DevBuf<unsigned char> rec;       // this rank's ArgminRecord + random-vector row
  DevBuf<unsigned char> rec_all;   // all-gathered records [world]
  bool nccl_owned = false;
  // best-agent exchange over peer memory (pmaf_p2p_export / pmaf_p2p_import); replaces the NCCL all-gather
  DevBuf<unsigned char> xchg;          // this rank's exchange block
  unsigned char *peer_xchg[kP2pMaxWorld] = {};
  bool p2p_ready = false;
  bool xchg_fresh = false;  // exported (zeroed) and not imported since: one import per export
  int p2p_rank = -1, p2p_world = 0;
  unsigned long long xseq = 0;
  DevBuf<unsigned long long> step_counter;
  DevBuf<HostOut> d_out;  // eval, best, real, real_path_out, step_counter are views into it
  DevBuf<unsigned char> l2_scratch;
  DevBuf<long long> section_cycles;
  DevBuf<unsigned> runtime_zero;
  // host mirrors
  std::vector<double> h_obs_pos, h_obs_vel, h_obs_rad;  // agents' copy
  std::vector<double> h_live;                            // last uploaded live list (pos|vel|rad)
  std::vector<double> real_path;
  std::vector<double> h_random;  // [A][O][3]
  RealState h_real{};
  DeviceBest h_best{0, 0, -1, 0};
  int best_n_obs = 0;
  EvalResult h_eval{};
  CostParams last_cost{};
  bool have_cost = false;
  bool fused_valid = false;
  ObstacleImage img{};
  float margin = 1e-3f;
  double margin_scale_static = 0.0;  // largest |coordinate| among goal, start, obstacles (broad_phase_margin)
  bool image_current = false;        // the staging image matches the agents' obstacle copy, layout and margin
  bool live_changed = false;         // upload_live sent a new list since the image was built
  int pending_feed_n = 0;            // pmaf_feed_obstacles not yet applied on the device
  double pending_feed_freq = 0.0;
  DevBuf<uint32_t> known_bits, known_keep;  // packed real-agent flags for the rollout's in-prologue reset
  // rollout bookkeeping
  bool rollout_pending = false;
  bool obstacles_advanced = false;  // a dynamic rollout advanced the agents' obstacle copies
  bool agents_touched = true;       // agent state changed since the last rollout launch
  int known_words = 0;
  // pinned staging
  double *h_stage = nullptr;
  size_t h_stage_doubles = 0;
  HostOut *h_out = nullptr;      // pinned + mapped
  HostOut *h_out_dev = nullptr;  // the same block as the device sees it
  unsigned long long ticket = 0;
  bool stage_busy = false;
  // rng for RandomCfAgent vectors
  bool seeded = false;
  uint64_t seed = 0;
  // tuning + counters
  int tune_lpa = 0, tune_block = 0, tune_occ = 0;
  pmaf_counters ctr{};
};

static void nccl_release(pmaf_planner *p);

static int set_device(pmaf_planner *p) {
  CU(cudaSetDevice(p->device));
  return 0;
}

#define ENTER(p)                                              \
  REQUIRE((p) != nullptr, PMAF_ERR_ARG, "null planner handle"); \
  if (int rc_ = set_device(p)) return rc_;

#define NEED_INIT(p) REQUIRE((p)->initialized, PMAF_ERR_STATE, "%s: pmaf_init has not been called", __func__)

static PlannerDev make_dev(const pmaf_planner *p) {
  PlannerDev d{};
  d.n_agents = p->A, d.first_agent = p->first_agent, d.n_obs = p->O, d.max_steps = p->H;
  for (int i = 0; i < 3; ++i) d.goal[i] = p->goal[i];
  d.shell = p->shell, d.mass = p->mass, d.rad = p->rad, d.vel_max = p->vel_max, d.approach_dist = p->approach;
  d.pred_dt = p->pred_dt;
  d.k_attr = p->k_attr.p, d.k_circ = p->k_circ.p, d.k_repel = p->k_repel.p, d.k_damp = p->k_damp.p;
  d.init_pos = p->init_pos.p, d.cur_pos = p->cur_pos.p, d.vel = p->vel.p, d.min_obs_dist = p->min_obs.p;
  d.path_len = p->path_len.p, d.ws_cost = p->ws_cost.p, d.pred_time_ns = p->pred_time.p;
  d.n_path = p->n_path.p, d.reached = p->reached.p, d.cost = p->cost.p, d.paths = p->paths.p;
  d.rot = p->rot.p, d.random_vecs = p->random_vecs.p, d.known = p->known.p, d.known_words = p->known_words;
  d.image = p->image.p, d.img = p->img;
  d.fused_cost = p->last_cost, d.fused_valid = p->fused_valid ? 1 : 0;
  d.step_counter = p->step_counter.p;
  d.section_cycles = p->section_cycles.p;
  d.runtime_zero = p->runtime_zero.p;
  d.reset_in_prologue = 0, d.reset_real = p->real.p;
  d.reset_known_bits = p->known_bits.p, d.reset_known_keep = p->known_keep.p;
  return d;
}

static ObstacleImage layout_image(int n_obs, bool dynamic) {
  ObstacleImage im{};
  im.n_obs = n_obs, im.dynamic = dynamic ? 1 : 0;
  const uint32_t col = (uint32_t)(((size_t)n_obs * sizeof(double) + 15) & ~(size_t)15);
  uint32_t off = 0;
  im.off_px = off, off += col;
  im.off_py = off, off += col;
  im.off_pz = off, off += col;
  im.off_rs = off, off += col;
  if (dynamic) {
    im.off_vx = off, off += col;
    im.off_vy = off, off += col;
    im.off_vz = off, off += col;
    im.off_dx = off, off += col;
    im.off_dy = off, off += col;
    im.off_dz = off, off += col;
  }
  im.off_bp = off, off += (uint32_t)((size_t)n_obs * sizeof(float4));
  im.off_nn = off, off += (uint32_t)(((size_t)n_obs * sizeof(uint16_t) + 15) & ~(size_t)15);
  im.nn_valid = 0;
  im.bytes = (off + 15u) & ~15u;
  return im;
}

// The OBSTACLE / GOAL_OBSTACLE agents (global indices 2 and 3) look up the nearest other obstacle of every
// obstacle they detect; in a static scene that is a per-tick table, built by spare blocks of reset_kernel.
static int wants_nn_table(const pmaf_planner *p, const ObstacleImage &im) {
  return (!im.dynamic && p->O >= 3 && p->first_agent <= 3 && p->first_agent + p->A > 2) ? 1 : 0;
}

static bool any_nonzero(const std::vector<double> &v) {
  for (double x : v)
    if (x != 0.0) return true;
  return false;
}

// Broad-phase safety margin (metres): dominates the fp32 rounding of coordinates up to `s` in magnitude.
// s = the largest coordinate the rollout can meet: goal, start, obstacles (static part, recomputed when the
// obstacle list changes) and the real agent's position, plus the farthest an agent / obstacle can travel.
static void refresh_margin_scale(pmaf_planner *p) {
  double s = 0.0;
  for (int i = 0; i < 3; ++i) s = std::max({s, std::fabs(p->goal[i]), std::fabs(p->mgr_init_pos[i])});
  for (double x : p->h_obs_pos) s = std::max(s, std::fabs(x));
  double vmax_obs = 0.0;
  for (double x : p->h_obs_vel) vmax_obs = std::max(vmax_obs, std::fabs(x));
  s += (p->vel_max + vmax_obs * 1.7320508) * p->pred_dt * p->H;
  p->margin_scale_static = s;
}
static float broad_phase_margin(const pmaf_planner *p, const double *pos) {
  double s = p->margin_scale_static;
  for (int i = 0; i < 3; ++i) s = std::max(s, std::fabs(pos[i]) + p->vel_max * p->pred_dt * p->H);
  if (!std::isfinite(s)) s = 1e6;
  return (float)(1e-3 + 4e-6 * s);
}
// the margin only ever grows between two inits (a larger margin is always safe), so that the staging image of
// a static scene stays valid from tick to tick
static void update_margin(pmaf_planner *p, const double *pos) {
  const float m = broad_phase_margin(p, pos);
  if (m > p->margin) p->margin = m, p->image_current = false;
}

// wait until the pinned H2D staging buffer may be overwritten
static int stage_acquire(pmaf_planner *p, size_t doubles) {
  if (p->stage_busy) {
    CU(cudaEventSynchronize(p->ev_stage));
    p->stage_busy = false;
  }
  if (doubles > p->h_stage_doubles) {
    if (p->h_stage) cudaFreeHost(p->h_stage);
    p->h_stage = nullptr, p->h_stage_doubles = 0;
    CU(cudaMallocHost(&p->h_stage, doubles * sizeof(double)));
    p->h_stage_doubles = doubles;
  }
  return 0;
}
static int stage_release(pmaf_planner *p) {
  CU(cudaEventRecord(p->ev_stage, p->stream));
  p->stage_busy = true;
  return 0;
}

static int h2d(pmaf_planner *p, void *dst, const void *src_pinned, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src_pinned, bytes, cudaMemcpyHostToDevice, p->stream));
  p->ctr.h2d_bytes += bytes;
  return 0;
}
static int d2h(pmaf_planner *p, void *dst, const void *src, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, p->stream));
  p->ctr.d2h_bytes += bytes;
  return 0;
}

// collect the device time of finished rollouts (CUDA events recorded around the kernel)
static int harvest_timing(pmaf_planner *p, int slot, bool wait) {
  if (!p->roll_timing[slot]) return 0;
  if (wait) CU(cudaEventSynchronize(p->ev_roll[slot][1]));
  else if (cudaEventQuery(p->ev_roll[slot][1]) != cudaSuccess) return 0;
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, p->ev_roll[slot][0], p->ev_roll[slot][1]));
  p->ctr.last_rollout_ms = ms;
  p->ctr.rollout_ms_total += ms;
  p->roll_timing[slot] = false;
  return 0;
}

static int finish_rollout(pmaf_planner *p) {
  if (p->rollout_pending) {
    CU(cudaStreamSynchronize(p->stream));
    p->rollout_pending = false;
  }
  if (int rc = harvest_timing(p, p->roll_slot ^ 1, true)) return rc;
  return harvest_timing(p, p->roll_slot, true);
}

// wait for the ticket a small kernel publishes in the mapped host block after storing its results there
static int wait_ticket(pmaf_planner *p, int which, unsigned long long ticket) {
  volatile unsigned long long *flag = &p->h_out->seq[which];
  timespec t0{};
  bool timed = false;
  long next_check_us = 2000;  // a failed launch or a trapped kernel never publishes: ask the stream, but rarely
  for (unsigned spins = 1; *flag != ticket; ++spins) {
    if ((spins & 0xfffu) != 0u) continue;
    timespec tn;
    clock_gettime(CLOCK_MONOTONIC, &tn);
    if (!timed) {
      t0 = tn, timed = true;
      continue;
    }
    const long us = (long)(tn.tv_sec - t0.tv_sec) * 1000000L + (tn.tv_nsec - t0.tv_nsec) / 1000L;
    if (us < next_check_us) continue;
    next_check_us = us + 2000;
    const cudaError_t e = cudaStreamQuery(p->stream);
    if (e == cudaErrorNotReady) continue;
    if (e != cudaSuccess) return fail(PMAF_ERR_CUDA, "stream failed while waiting for a result: %s", cudaGetErrorString(e));
    if (*flag != ticket) return fail(PMAF_ERR_CUDA, "a result ticket never arrived although the stream is idle");
  }
  __sync_synchronize();
  return 0;
}

// ---- NCCL (dlopen) ----------------------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  REQUIRE(g_nccl.handle != nullptr, PMAF_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.handle, "ncclAllGather");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
  REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather, PMAF_ERR_NCCL,
          "libnccl lacks a required symbol");
  return 0;
}

#define NC(call)                                                                                       \
  do {                                                                                                 \
    ncclResult_t r_ = (call);                                                                          \
    if (r_ != ncclSuccess)                                                                             \
      return fail(PMAF_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
  } while (0)

static void nccl_release(pmaf_planner *p) {
  if (p->nccl && p->nccl_owned && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)p->nccl);
  p->nccl = nullptr, p->nccl_owned = false;
}

extern "C" int pmaf_nccl_unique_id(unsigned char out[128]) {
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  if (int rc = nccl_load()) return rc;
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out, &id, 128);
  return 0;
}

extern "C" int pmaf_nccl_init(pmaf_planner *p, const unsigned char id_bytes[128], int rank, int world) {
  ENTER(p);
  REQUIRE(id_bytes && world >= 1 && rank >= 0 && rank < world, PMAF_ERR_ARG, "pmaf_nccl_init: bad argument");
  if (int rc = nccl_load()) return rc;
  nccl_release(p);
  ncclUniqueId id;
  memcpy(&id, id_bytes, 128);
  ncclComm_t comm = nullptr;
  NC(g_nccl.CommInitRank(&comm, world, id, rank));
  p->nccl = comm, p->nccl_owned = true;
  return 0;
}

// ---- best-agent exchange over peer memory (cudaIpc) --------------------------------------------------------------------
static void p2p_release(pmaf_planner *p) {
  for (int r = 0; r < kP2pMaxWorld; ++r) {
    if (p->peer_xchg[r] && p->peer_xchg[r] != p->xchg.p) cudaIpcCloseMemHandle(p->peer_xchg[r]);
    p->peer_xchg[r] = nullptr;
  }
  p->p2p_ready = false;
}

extern "C" int pmaf_p2p_export(pmaf_planner *p, unsigned char handle_out[64]) {
  ENTER(p);
  REQUIRE(handle_out, PMAF_ERR_ARG, "null output");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  p2p_release(p);
  CU(p->xchg.resize(p2p_block_bytes()));
  CU(cudaMemsetAsync(p->xchg.p, 0, p2p_block_bytes(), p->stream));
  CU(cudaStreamSynchronize(p->stream));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, p->xchg.p));
  memcpy(handle_out, &h, 64);
  p->xchg_fresh = true;
  return 0;
}

extern "C" int pmaf_p2p_import(pmaf_planner *p, const unsigned char *handles, int rank, int world) {
  ENTER(p);
  REQUIRE(handles && world >= 2 && world <= kP2pMaxWorld && rank >= 0 && rank < world, PMAF_ERR_ARG,
          "pmaf_p2p_import: bad argument (2 <= world <= %d)", kP2pMaxWorld);
  // one import per export: the export zeroes the block's sequence numbers (a second import would restart the
  // sequence against stale flags) and drops earlier mappings (a second import would leak them)
  REQUIRE(p->xchg.p != nullptr && p->xchg_fresh, PMAF_ERR_STATE,
          "pmaf_p2p_import: call pmaf_p2p_export first (every import needs a fresh export on all ranks)");
  p->xchg_fresh = false;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      p->peer_xchg[r] = p->xchg.p;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void *ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      p2p_release(p);
      (void)cudaGetLastError();
      return fail(PMAF_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
    p->peer_xchg[r] = static_cast<unsigned char *>(ptr);
  }
  p->h_out->p2p_fail = 0;
  p->xseq = 0;
  p->p2p_rank = rank, p->p2p_world = world;
  p->p2p_ready = true;
  return 0;
}

// ---- lifecycle -------------------------------------------------------------------------------------------------
static int create_resources(pmaf_planner *p) {
  CU(cudaSetDevice(p->device));
  CU(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  for (int s = 0; s < 2; ++s)
    for (int e = 0; e < 2; ++e) CU(cudaEventCreate(&p->ev_roll[s][e]));
  CU(cudaEventCreateWithFlags(&p->ev_d2h, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&p->ev_stage, cudaEventDisableTiming));
  CU(cudaEventCreate(&p->ev_t0));
  CU(cudaEventCreate(&p->ev_t1));
  CU(cudaHostAlloc(&p->h_out, sizeof(HostOut), cudaHostAllocMapped));
  memset(p->h_out, 0, sizeof(HostOut));
  CU(cudaHostGetDevicePointer(&p->h_out_dev, p->h_out, 0));
  CU(p->d_out.resize(1));
  p->eval.alias(&p->d_out.p->eval, 1), p->best.alias(&p->d_out.p->best, 1), p->real.alias(&p->d_out.p->real, 1);
  p->real_path_out.alias(p->d_out.p->real_path, 256 * 3), p->step_counter.alias(p->d_out.p->steps, 16);
  CU(p->rec.resize(argmin_record_bytes(0)));
  CU(p->scratch.resize(16));
  CU(p->runtime_zero.resize(1));
  CU(cudaMemsetAsync(p->runtime_zero.p, 0, sizeof(unsigned), p->stream));
#if defined(PMAF_SECTION_TIMERS) || defined(PMAF_FAST_STATS)
  CU(p->section_cycles.resize(64 * 12));
  CU(cudaMemsetAsync(p->section_cycles.p, 0, 64 * 12 * sizeof(long long), p->stream));
#endif
  CU(cudaMemsetAsync(p->best.p, 0, sizeof(DeviceBest), p->stream));
  CU(cudaMemsetAsync(p->real.p, 0, sizeof(RealState), p->stream));
  CU(cudaMemsetAsync(p->step_counter.p, 0, 16 * sizeof(unsigned long long), p->stream));
  CU(cudaStreamSynchronize(p->stream));
  return 0;
}

extern "C" int pmaf_destroy(pmaf_planner *p);

extern "C" int pmaf_create(pmaf_planner **out, int device) {
  REQUIRE(out != nullptr, PMAF_ERR_ARG, "pmaf_create: out is null");
  *out = nullptr;
  int n_dev = 0;
  CU(cudaGetDeviceCount(&n_dev));
  REQUIRE(device >= 0 && device < n_dev, PMAF_ERR_CUDA, "pmaf_create: CUDA device %d not available (%d visible)",
          device, n_dev);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  // the library holds sm_100a code only, and arch-specific ("a") code runs on exactly that architecture
  REQUIRE(prop.major == 10 && prop.minor == 0, PMAF_ERR_CUDA,
          "pmaf_create: device %d is sm_%d%d; libpmaf is built for sm_100a only and has no fallback", device,
          prop.major, prop.minor);
  pmaf_planner *p = new (std::nothrow) pmaf_planner();
  REQUIRE(p != nullptr, PMAF_ERR_ALLOC, "pmaf_create: out of host memory");
  p->device = device;
  if (int rc = create_resources(p)) {
    const std::string why = g_last_error;  // pmaf_destroy must not hide the cause
    pmaf_destroy(p);
    g_last_error = why;
    return rc;
  }
  *out = p;
  return 0;
}

extern "C" int pmaf_destroy(pmaf_planner *p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  for (DevBuf<double> *b :
       {&p->k_attr, &p->k_circ, &p->k_repel, &p->k_damp, &p->init_pos, &p->cur_pos, &p->vel, &p->min_obs,
        &p->path_len, &p->ws_cost, &p->pred_time, &p->cost, &p->paths, &p->rot, &p->random_vecs, &p->obs_pos,
        &p->obs_vel, &p->obs_rad, &p->live_pos, &p->live_vel, &p->live_rad, &p->real_rot, &p->best_random,
        &p->scratch, &p->real_path_out})
    b->release();
  p->l2_scratch.release();
  p->known_bits.release(), p->known_keep.release();
  p->section_cycles.release();
  p->runtime_zero.release();
  p->n_path.release(), p->reached.release(), p->known.release(), p->image.release(), p->real_known.release();
  p->real.release(), p->best.release(), p->eval.release(), p->rec.release(), p->step_counter.release();
  p->rec_all.release();
  p->d_out.release();
  p2p_release(p);
  p->xchg.release();
  nccl_release(p);
  if (p->h_stage) cudaFreeHost(p->h_stage);
  if (p->h_out) cudaFreeHost(p->h_out);
  for (cudaEvent_t e : {p->ev_roll[0][0], p->ev_roll[0][1], p->ev_roll[1][0], p->ev_roll[1][1], p->ev_d2h, p->ev_stage, p->ev_t0, p->ev_t1})
    if (e) cudaEventDestroy(e);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return 0;
}

extern "C" int pmaf_set_shard(pmaf_planner *p, int n_global, int first_agent, int rank, int world) {
  ENTER(p);
  REQUIRE(n_global >= 1 && first_agent >= 0 && first_agent < n_global && world >= 1 && rank >= 0 && rank < world,
          PMAF_ERR_ARG, "pmaf_set_shard: bad shard (n_global=%d first=%d rank=%d world=%d)", n_global, first_agent,
          rank, world);
  p->n_global = n_global, p->first_agent = first_agent, p->rank = rank, p->world = world;
  p->shard_set = true;
  return 0;
}

extern "C" int pmaf_set_nccl_comm(pmaf_planner *p, void *nccl_comm) {
  ENTER(p);
  if (int rc = nccl_load()) return rc;
  nccl_release(p);
  p->nccl = nccl_comm, p->nccl_owned = false;
  return 0;
}

extern "C" int pmaf_seed_random_vecs(pmaf_planner *p, uint64_t seed) {
  ENTER(p);
  p->seeded = true, p->seed = seed;
  return 0;
}

static int upload_real(pmaf_planner *p) {
  if (int rc = stage_acquire(p, sizeof(RealState) / sizeof(double))) return rc;
  memcpy(p->h_stage, &p->h_real, sizeof(RealState));
  if (int rc = h2d(p, p->real.p, p->h_stage, sizeof(RealState))) return rc;
  return stage_release(p);
}

template <class K, class... Args>
static int launch(pmaf_planner *p, K kernel, dim3 grid, dim3 block, size_t smem, Args... args) {
  kernel<<<grid, block, smem, p->stream>>>(args...);
  CU(cudaGetLastError());
  p->ctr.kernel_launches++;
  return 0;
}
// Programmatic dependent launch: the grid may start while its predecessor in the stream (which executes
// griddepcontrol.launch_dependents) is still running; the kernel itself waits (griddepcontrol.wait) before it
// touches anything the predecessor writes. Hides the launch latency of the rollout behind tick_kernel.
template <class... KArgs, class... Args>
static int launch_dependent(pmaf_planner *p, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = p->stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr, cfg.numAttrs = 1;
  CU(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
  p->ctr.kernel_launches++;
  return 0;
}

// (re)build the staging image and, optionally, reset the agents
static int launch_reset(pmaf_planner *p, bool do_agents, bool from_real, const double *dev_pos_vel, int n_update,
                        const double *dev_new_pos, const double *dev_new_vel, bool set_known, bool reset_velocity) {
  PlannerDev d = make_dev(p);
  ResetArgs r{};
  r.pos_vel = dev_pos_vel, r.real = p->real.p, r.from_real = from_real ? 1 : 0;
  r.n_obs_update = n_update, r.new_pos = dev_new_pos, r.new_vel = dev_new_vel;
  r.obs_pos = p->obs_pos.p, r.obs_vel = p->obs_vel.p, r.obs_rad = p->obs_rad.p;
  r.real_known = p->real_known.p, r.image = p->image.p, r.margin = p->margin;
  r.set_known = set_known ? 1 : 0, r.reset_velocity = reset_velocity ? 1 : 0;
  r.do_agents = do_agents ? 1 : 0;
  const int block = 128;
  r.agent_blocks = do_agents ? (p->A + block - 1) / block : 1;
  // nearest-neighbour table: one warp per obstacle row, spread over extra blocks (they overlap block 0)
  const int nn_blocks = p->img.nn_valid ? std::min(64, (p->O - 1 + 3) / 4) : 0;
  p->image_current = true, p->live_changed = false;
  return launch(p, reset_kernel, dim3(r.agent_blocks + nn_blocks), dim3(block), 0, d, r);
}

extern "C" int pmaf_init(pmaf_planner *p, const double goal[3], double delta_t, int n_obs, const double *obs_pos,
                         const double *obs_vel, const double *obs_rad, int n_agents, const double *k_attr,
                         const double *k_circ, const double *k_repel, const double *k_damp, const double *k_manip,
                         int n_force, const double *k_repel_force, double velocity_max, double approach_dist,
                         double detect_shell_rad, uint64_t max_prediction_steps, uint64_t prediction_freq_multiple,
                         double agent_mass, double radius) {
  ENTER(p);
  (void)k_manip, (void)n_force, (void)k_repel_force;  // manipulability / body forces have no caller on this path
  REQUIRE(goal && obs_pos && obs_vel && obs_rad && k_attr && k_circ && k_repel && k_damp, PMAF_ERR_ARG,
          "pmaf_init: null array");
  REQUIRE(n_obs >= 1 && n_obs <= kMaxObstacles, PMAF_ERR_ARG,
          "pmaf_init: n_obs=%d out of range [1, %d] (the last obstacle is the sentinel)", n_obs, kMaxObstacles);
  REQUIRE(n_agents >= 0, PMAF_ERR_ARG, "pmaf_init: n_agents=%d", n_agents);
  REQUIRE(max_prediction_steps >= 1 && max_prediction_steps <= (1u << 24), PMAF_ERR_ARG,
          "pmaf_init: max_prediction_steps=%llu out of range", (unsigned long long)max_prediction_steps);
  if (int rc = finish_rollout(p)) return rc;
  CU(cudaStreamSynchronize(p->stream));
  p->initialized = false;  // a failure below leaves the planner uninitialised, not half re-initialised

  // at least the HAD agent always exists (cf_manager.cpp:70-72); gains of a missing agent read as 0
  const int n_glob_in = std::max(n_agents, 1);
  if (!p->shard_set) p->n_global = n_glob_in, p->first_agent = 0, p->rank = 0, p->world = 1;
  REQUIRE(p->n_global == n_glob_in, PMAF_ERR_ARG,
          "pmaf_init: sharded planner expects the GLOBAL gain arrays (n_agents=%d, shard n_global=%d)", n_agents,
          p->n_global);
  int n_local = p->n_global;
  if (p->shard_set) {
    // contiguous blocks: rank r owns [r*n/w, (r+1)*n/w); first_agent was given by the caller
    long long next = (long long)(p->rank + 1) * p->n_global / p->world;
    n_local = (int)(next - p->first_agent);
    REQUIRE(n_local >= 1, PMAF_ERR_ARG, "pmaf_init: shard of rank %d is empty", p->rank);
  }
  p->A = n_local, p->O = n_obs, p->H = (int)max_prediction_steps;
  for (int i = 0; i < 3; ++i) p->goal[i] = goal[i];
  p->delta_t = delta_t, p->pred_dt = (double)prediction_freq_multiple * delta_t;
  p->shell = detect_shell_rad, p->mass = agent_mass, p->rad = radius, p->vel_max = velocity_max;
  p->approach = approach_dist;
  p->known_words = (n_obs + 31) / 32;
  const size_t A = p->A, O = p->O, H = p->H, G = p->n_global;

  CU(p->k_attr.resize(G));
  CU(p->k_circ.resize(G));
  CU(p->k_repel.resize(G));
  CU(p->k_damp.resize(G));
  CU(p->init_pos.resize(A * 3));
  CU(p->cur_pos.resize(A * 3));
  CU(p->vel.resize(A * 3));
  CU(p->min_obs.resize(A));
  CU(p->path_len.resize(A));
  CU(p->ws_cost.resize(A));
  CU(p->pred_time.resize(A));
  CU(p->cost.resize(A));
  CU(p->n_path.resize(A));
  CU(p->reached.resize(A));
  CU(p->paths.resize(A * H * 3));
  CU(p->rot.resize(A * O * 3));
  CU(p->random_vecs.resize(A * O * 3));
  CU(p->known.resize(A * p->known_words));
  CU(p->obs_pos.resize(O * 3));
  CU(p->obs_vel.resize(O * 3));
  CU(p->obs_rad.resize(O));
  CU(p->live_pos.resize(O * 7));  // pos | vel | rad in one block: one upload per call
  p->live_vel.alias(p->live_pos.p + 3 * O, 3 * O), p->live_rad.alias(p->live_pos.p + 6 * O, O);
  CU(p->real_known.resize(O));
  CU(p->known_bits.resize(p->known_words));
  CU(p->known_keep.resize(p->known_words));
  CU(p->real_rot.resize(O * 3));
  CU(p->rec.resize(argmin_record_bytes((int)O)));
  CU(p->rec_all.resize(argmin_record_bytes((int)O) * (size_t)p->world));
  {  // best_agent_ survives init (quirk 6); keep the overlapping part of its random vectors
    DevBuf<double> old = p->best_random;
    const int old_n = p->best_n_obs;
    p->best_random = DevBuf<double>();
    CU(p->best_random.resize(O * 3));
    CU(cudaMemsetAsync(p->best_random.p, 0, O * 3 * sizeof(double), p->stream));
    if (old.p && old_n > 0)
      CU(cudaMemcpyAsync(p->best_random.p, old.p, (size_t)std::min<int>(old_n, (int)O) * 3 * sizeof(double),
                         cudaMemcpyDeviceToDevice, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    old.release();
    p->best_n_obs = (int)O;
  }
  p->h_live.clear();
  p->pending_feed_n = 0;  // a feed nobody consumed belongs to the list of the previous init

  // gains (global), agents' obstacle copy, random vectors
  p->h_obs_pos.assign(obs_pos, obs_pos + O * 3);
  p->h_obs_vel.assign(obs_vel, obs_vel + O * 3);
  p->h_obs_rad.assign(obs_rad, obs_rad + O);
  {
    const size_t n = 4 * G + 7 * O;
    if (int rc = stage_acquire(p, n)) return rc;
    double *s = p->h_stage;
    for (size_t i = 0; i < G; ++i) {
      const bool have = (int)i < n_agents;
      s[i] = have ? k_attr[i] : 0.0, s[G + i] = have ? k_circ[i] : 0.0;
      s[2 * G + i] = have ? k_repel[i] : 0.0, s[3 * G + i] = have ? k_damp[i] : 0.0;
    }
    memcpy(s + 4 * G, obs_pos, O * 3 * sizeof(double));
    memcpy(s + 4 * G + 3 * O, obs_vel, O * 3 * sizeof(double));
    memcpy(s + 4 * G + 6 * O, obs_rad, O * sizeof(double));
    if (int rc = h2d(p, p->k_attr.p, s, G * sizeof(double))) return rc;
    if (int rc = h2d(p, p->k_circ.p, s + G, G * sizeof(double))) return rc;
    if (int rc = h2d(p, p->k_repel.p, s + 2 * G, G * sizeof(double))) return rc;
    if (int rc = h2d(p, p->k_damp.p, s + 3 * G, G * sizeof(double))) return rc;
    if (int rc = h2d(p, p->obs_pos.p, s + 4 * G, O * 3 * sizeof(double))) return rc;
    if (int rc = h2d(p, p->obs_vel.p, s + 4 * G + 3 * O, O * 3 * sizeof(double))) return rc;
    if (int rc = h2d(p, p->obs_rad.p, s + 4 * G + 6 * O, O * sizeof(double))) return rc;
    if (int rc = stage_release(p)) return rc;
  }
  {  // RandomCfAgent vectors (cf_agent.h:338-342): U[-1,1]^3 normalised, agent-major, obstacle-minor.
     // The reference seeds a fresh mt19937 from std::random_device per vector
     // (helper_functions.cpp:8-12); one generator seeded once gives the same distribution.
    p->h_random.assign(A * O * 3, 0.0);
    std::mt19937_64 gen(p->seeded ? p->seed : ((uint64_t)std::random_device{}() << 32) ^ std::random_device{}());
    std::uniform_real_distribution<double> dis(-1.0, 1.0);
    if (p->seeded) {  // keep streams shard-independent: skip the draws of agents before this shard
      const unsigned long long skip_agents = p->first_agent > 5 ? p->first_agent - 5 : 0;
      gen.discard(skip_agents * O * 3);
    }
    for (size_t a = 0; a < A; ++a) {
      if (agent_type_of_index(p->first_agent + (int)a) != RANDOM_AGENT) continue;
      for (size_t i = 0; i < O; ++i) {
        v3 r = normalized3(mk3(dis(gen), dis(gen), dis(gen)));
        st3(&p->h_random[(a * O + i) * 3], r);
      }
    }
    CU(cudaMemcpyAsync(p->random_vecs.p, p->h_random.data(), A * O * 3 * sizeof(double), cudaMemcpyHostToDevice,
                       p->stream));
    p->ctr.h2d_bytes += A * O * 3 * sizeof(double);
    CU(cudaStreamSynchronize(p->stream));
  }

  // real agent re-created (cf_manager.cpp:66-68): pos = [init_pos_], vel = (0.01,0,0), init_pos_ = 0
  memset(&p->h_real, 0, sizeof p->h_real);
  for (int i = 0; i < 3; ++i) p->h_real.pos[i] = p->mgr_init_pos[i];
  p->h_real.vel[0] = 0.01;
  p->real_path.assign(p->mgr_init_pos, p->mgr_init_pos + 3);
  if (int rc = upload_real(p)) return rc;
  CU(cudaMemsetAsync(p->real_known.p, 0, O, p->stream));
  {
    std::vector<double> rr(O * 3, 0.0);
    for (size_t i = 0; i < O; ++i) rr[3 * i + 2] = 1.0;
    CU(cudaMemcpyAsync(p->real_rot.p, rr.data(), O * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CU(cudaStreamSynchronize(p->stream));
  }

  p->initialized = true;
  p->fused_valid = false;
  p->obstacles_advanced = false;
  p->agents_touched = true;
  p->img = layout_image((int)O, any_nonzero(p->h_obs_vel));
  p->img.nn_valid = wants_nn_table(p, p->img);
  CU(p->image.resize(p->img.bytes));
  refresh_margin_scale(p);
  p->margin = broad_phase_margin(p, p->mgr_init_pos);
  p->image_current = false, p->live_changed = false;

  // agents constructed at the manager's init_pos_ (cf_manager.cpp:70-104)
  {
    if (int rc = stage_acquire(p, 3)) return rc;
    memcpy(p->h_stage, p->mgr_init_pos, 3 * sizeof(double));
    if (int rc = h2d(p, p->scratch.p, p->h_stage, 3 * sizeof(double))) return rc;
    if (int rc = stage_release(p)) return rc;
    PlannerDev d = make_dev(p);
    if (int rc = launch(p, init_state_kernel, dim3(296), dim3(256), 0, d, (const double *)p->scratch.p)) return rc;
  }
  if (int rc = launch_reset(p, false, false, nullptr, 0, nullptr, nullptr, false, false)) return rc;
  CU(cudaStreamSynchronize(p->stream));
  return 0;
}

extern "C" int pmaf_set_random_vecs(pmaf_planner *p, const double *vecs, int n_agents, int n_obs) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(vecs != nullptr, PMAF_ERR_ARG, "pmaf_set_random_vecs: null array");
  REQUIRE(n_obs == p->O && (n_agents == p->A || n_agents == p->n_global), PMAF_ERR_ARG,
          "pmaf_set_random_vecs: expected [%d or %d][%d][3], got [%d][%d][3]", p->A, p->n_global, p->O, n_agents,
          n_obs);
  if (int rc = finish_rollout(p)) return rc;
  const size_t O = p->O;
  const size_t off = n_agents == p->A ? 0 : (size_t)p->first_agent;  // local table or slice of the global one
  for (size_t a = 0; a < (size_t)p->A; ++a) {
    if (agent_type_of_index(p->first_agent + (int)a) != RANDOM_AGENT) continue;
    memcpy(&p->h_random[a * O * 3], vecs + (off + a) * O * 3, O * 3 * sizeof(double));
  }
  CU(cudaMemcpyAsync(p->random_vecs.p, p->h_random.data(), p->h_random.size() * sizeof(double),
                     cudaMemcpyHostToDevice, p->stream));
  p->ctr.h2d_bytes += p->h_random.size() * sizeof(double);
  CU(cudaStreamSynchronize(p->stream));
  return 0;
}

extern "C" int pmaf_get_random_vecs(pmaf_planner *p, double *vecs, int n_agents, int n_obs) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(vecs && n_obs == p->O && n_agents == p->A, PMAF_ERR_ARG, "pmaf_get_random_vecs: expected [%d][%d][3]",
          p->A, p->O);
  memcpy(vecs, p->h_random.data(), p->h_random.size() * sizeof(double));
  return 0;
}

// ---- per-tick calls ----------------------------------------------------------------------------------------------------
extern "C" int pmaf_set_initial_position(pmaf_planner *p, const double pos[3]) {
  ENTER(p);
  REQUIRE(pos != nullptr, PMAF_ERR_ARG, "pmaf_set_initial_position: null position");
  for (int i = 0; i < 3; ++i) p->mgr_init_pos[i] = pos[i];
  if (!p->initialized) return 0;  // default-constructed manager: only init_pos_ is meaningful
  if (int rc = finish_rollout(p)) return rc;
  // real agent: init_pos_ := pos and the path is APPENDED to (cf_agent.cpp:34-46)
  for (int i = 0; i < 3; ++i) p->h_real.init_pos[i] = pos[i], p->h_real.pos[i] = pos[i];
  p->real_path.insert(p->real_path.end(), pos, pos + 3);
  if (int rc = upload_real(p)) return rc;
  if (int rc = stage_acquire(p, 3)) return rc;
  memcpy(p->h_stage, pos, 3 * sizeof(double));
  if (int rc = h2d(p, p->scratch.p, p->h_stage, 3 * sizeof(double))) return rc;
  if (int rc = stage_release(p)) return rc;
  PlannerDev d = make_dev(p);
  if (int rc = launch(p, set_initial_position_kernel, dim3((p->A + 127) / 128), dim3(128), 0, d,
                      (const double *)p->scratch.p))
    return rc;
  p->fused_valid = false;
  p->agents_touched = true;
  return 0;
}

extern "C" int pmaf_set_real_position(pmaf_planner *p, const double pos[3]) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(pos != nullptr, PMAF_ERR_ARG, "pmaf_set_real_position: null position");
  for (int i = 0; i < 3; ++i) p->h_real.pos[i] = pos[i];
  p->real_path.insert(p->real_path.end(), pos, pos + 3);
  return upload_real(p);
}

template <int LPA>
static int launch_rollout_lpa(pmaf_planner *p, const PlannerDev &d, int block, bool dynamic, bool dependent) {
  const int groups = block / LPA;
  const int grid = (p->A + groups - 1) / groups;
  const size_t smem = rollout_smem_bytes(p->img, groups, LPA, p->known_words);
  REQUIRE(smem <= 227 * 1024, PMAF_ERR_ARG, "rollout needs %zu B of shared memory (> 227 KB): too many obstacles",
          smem);
  // many warps per SM: the 128-register build keeps 16 warps resident; few: the 255-register build
  int occ = p->tune_occ;
  // measured on B200: the 255-register latency build wins up to ~8 warps per SM (C5: 1024 agents), the
  // 170-register build beyond (C3: 4096, C4 shard: 8192 agents) — unless the scene is static with more than
  // 64 field obstacles, where the latency build's chunked straight-line step (MULTI) is used
  const bool multi_ok = LPA == 32 && !dynamic && p->O - 1 > kBroadUnrolledRounds * 32 && block <= 256;
  if (occ == 0) occ = (LPA == 32 && block <= 128 && (long long)grid * block / 32 > 12 * 148 && !multi_ok) ? 3 : 1;
  if (LPA < 8 || block > 128) occ = 1;  // the occupancy builds exist for 8/16/32 lanes per agent, CTAs <= 128 threads
  auto kern = dynamic ? rollout_kernel<LPA, true, 1> : rollout_kernel<LPA, false, 1>;
  if constexpr (LPA == 32) {
    if (occ == 1 && multi_ok) kern = rollout_kernel<32, false, 1, true>;
  }
  if constexpr (LPA >= 8) {
    if (occ == 3) kern = dynamic ? rollout_kernel<LPA, true, 3> : rollout_kernel<LPA, false, 3>;

    if (occ == 4) kern = dynamic ? rollout_kernel<LPA, true, 4> : rollout_kernel<LPA, false, 4>;
  }
  p->ctr.occupancy_build = occ;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p->ctr.lanes_per_agent = LPA, p->ctr.block_threads = block, p->ctr.grid_blocks = grid, p->ctr.smem_bytes = (int)smem;
  if (dependent) return launch_dependent(p, kern, dim3(grid), dim3(block), smem, d);
  return launch(p, kern, dim3(grid), dim3(block), smem, d);
}

static void pick_rollout_shape(const pmaf_planner *p, int &lpa, int &block) {
  // One warp per agent (32 lanes) while the population fits the machine about twice over (8 resident warps per SM
  // at 255 registers: 1184 agents in flight); beyond that 4 agents share a warp (8 lanes each, fast_step_packed):
  // measured on B200, 4096 agents x 256 obstacles 6.8 -> 5.0 ms, 8192 x 1024 15.3 -> 10.9 ms, while 1024 agents
  // x 50 moving obstacles lose (0.65 -> 0.78 ms: a packed warp's step is a longer dependency chain).
  lpa = p->tune_lpa ? p->tune_lpa : (p->A >= 3072 ? 8 : 32);
  if (p->tune_block) {
    block = p->tune_block;
  } else {
    // small populations are latency-bound: one warp per CTA spreads them over all 148 SMs; large ones pack
    // 2 warps per CTA (measured best on 4096 agents x 256 obstacles with one warp per agent), packed shapes a
    // whole SM's worth (8 warps share one obstacle image)
    const long long warps = ((long long)p->A * lpa + 31) / 32;
    block = warps <= 2 * 148 ? 32 : (lpa < 32 && warps >= 8 * 128 ? 256 : 64);
    // the obstacle image plus the per-agent lists must fit 227 KB
    while (block > 32 && block > lpa && rollout_smem_bytes(p->img, block / lpa, lpa, p->known_words) > 227 * 1024) block /= 2;
  }
  if (block < lpa) block = lpa;
}

static int launch_rollout(pmaf_planner *p, bool reset_in_prologue = false) {
  PlannerDev d = make_dev(p);
  d.reset_in_prologue = reset_in_prologue ? 1 : 0;
  int lpa, block;
  pick_rollout_shape(p, lpa, block);
  // per-rollout step counter: zeroed by tick_kernel on the fused path (one stream operation less per tick)
  if (!reset_in_prologue) CU(cudaMemsetAsync(p->step_counter.p, 0, sizeof(unsigned long long), p->stream));
  const bool timed = p->time_rollouts;
  // untimed rollouts on the fused path start as programmatic dependents of tick_kernel (an event record in
  // between would serialise the two again)
  const bool dependent = reset_in_prologue && !timed;
  if (timed) {
    p->roll_slot ^= 1;  // the slot used two rollouts ago: finished long before this point in the stream
    if (int rc = harvest_timing(p, p->roll_slot, true)) return rc;
    CU(cudaEventRecord(p->ev_roll[p->roll_slot][0], p->stream));
  }
  int rc;
  const bool dyn = p->img.dynamic != 0;
  switch (lpa) {
    case 4: rc = launch_rollout_lpa<4>(p, d, block, dyn, dependent); break;
    case 8: rc = launch_rollout_lpa<8>(p, d, block, dyn, dependent); break;
    case 16: rc = launch_rollout_lpa<16>(p, d, block, dyn, dependent); break;
    default: rc = launch_rollout_lpa<32>(p, d, block, dyn, dependent); break;
  }
  if (rc) return rc;
  if (timed) {
    CU(cudaEventRecord(p->ev_roll[p->roll_slot][1], p->stream));
    p->roll_timing[p->roll_slot] = true;
  }
  p->ctr.rollouts++;
  p->rollout_pending = true;
  if (dyn) p->obstacles_advanced = true;
  p->agents_touched = false;
  return 0;
}

extern "C" int pmaf_start_prediction(pmaf_planner *p) {
  ENTER(p);
  NEED_INIT(p);
  if (!p->agents_touched) return 0;  // every agent's stop condition already holds: nothing to run
  REQUIRE(!(p->img.dynamic && p->obstacles_advanced), PMAF_ERR_STATE,
          "pmaf_start_prediction: agents were repositioned without pmaf_reset_agents after a rollout over moving "
          "obstacles; continuing from per-agent advanced obstacle copies is not supported");
  if (int rc = finish_rollout(p)) return rc;
  return launch_rollout(p);
}

extern "C" int pmaf_stop_prediction(pmaf_planner *p) {
  ENTER(p);
  return finish_rollout(p);
}

static bool same_cost(const CostParams &a, const CostParams &b) { return memcmp(&a, &b, sizeof a) == 0; }

static int launch_evaluate(pmaf_planner *p, const CostParams &C) {
  PlannerDev d = make_dev(p);
  if (!(p->fused_valid && p->have_cost && same_cost(C, p->last_cost))) {
    if (int rc = launch(p, workspace_cost_kernel, dim3((p->A + 127) / 128), dim3(128), 0, d, C)) return rc;
  }
  p->last_cost = C, p->have_cost = true;
  const int threads = p->A >= 1024 ? 1024 : std::max(32, ((p->A + 31) / 32) * 32);
  ArgminRecord *rec = reinterpret_cast<ArgminRecord *>(p->rec.p);
  if (p->world == 1) {
    p->ctr.d2h_bytes += sizeof(EvalResult) + sizeof(DeviceBest);  // stored by the kernel into the mapped host block
    return launch(p, evaluate_kernel, dim3(1), dim3(threads), 0, d, C, p->best.p, p->best_random.p, rec, p->eval.p, 1,
                  p->h_out_dev, ++p->ticket);
  }
  // sharded: local scan, then the exchange + replicated selection
  if (int rc = launch(p, evaluate_kernel, dim3(1), dim3(threads), 0, d, C, p->best.p, p->best_random.p, rec,
                      p->eval.p, 0, (HostOut *)nullptr, 0ull))
    return rc;
  if (p->p2p_ready) {  // ONE kernel: P2P stores into every peer's block over NVLink, wait, select, publish to the host
    REQUIRE(p->p2p_rank == p->rank && p->p2p_world == p->world, PMAF_ERR_STATE,
            "peer-memory exchange was set up for rank %d of %d, the shard is rank %d of %d", p->p2p_rank, p->p2p_world,
            p->rank, p->world);
    P2pExchange X{};
    for (int r = 0; r < p->world; ++r) X.peers[r] = p->peer_xchg[r];
    X.rank = p->rank, X.world = p->world, X.seq = ++p->xseq, X.stride = p2p_slot_stride();
    p->ctr.collectives++;
    p->ctr.d2h_bytes += sizeof(EvalResult) + sizeof(DeviceBest);
    return launch(p, p2p_select_kernel, dim3(1), dim3(256), 0, (const unsigned char *)p->rec.p, X, p->O, p->best.p,
                  p->best_random.p, p->eval.p, p->h_out_dev, ++p->ticket, (int *)&p->h_out_dev->p2p_fail);
  }
  REQUIRE(p->nccl != nullptr, PMAF_ERR_STATE, "sharded planner without an exchange (pmaf_nccl_init or pmaf_p2p_import)");
  const size_t bytes = argmin_record_bytes(p->O);
  NC(g_nccl.AllGather(p->rec.p, p->rec_all.p, bytes, ncclChar, (ncclComm_t)p->nccl, p->stream));
  p->ctr.collectives++;
  return launch(p, global_select_kernel, dim3(1), dim3(256), 0, (const unsigned char *)p->rec_all.p, p->world, p->O,
                p->best.p, p->best_random.p, p->eval.p);
}

static CostParams make_cost(double k_goal_dist, double k_path_len, double k_safe_dist, double k_workspace,
                            const double ws[6]) {
  CostParams C{};
  C.k_goal_dist = k_goal_dist, C.k_path_len = k_path_len, C.k_safe_dist = k_safe_dist, C.k_workspace = k_workspace;
  for (int i = 0; i < 6; ++i) C.ws[i] = ws[i];
  return C;
}

extern "C" int pmaf_evaluate_agents(pmaf_planner *p, int n_obs, const double *obs_pos, const double *obs_vel,
                                    const double *obs_rad, double k_goal_dist, double k_path_len,
                                    double k_safe_dist, double k_workspace, const double ws_limits[6],
                                    int *best_index) {
  ENTER(p);
  NEED_INIT(p);
  (void)n_obs, (void)obs_pos, (void)obs_vel, (void)obs_rad;  // unused by the reference as well
  REQUIRE(ws_limits && best_index, PMAF_ERR_ARG, "pmaf_evaluate_agents: null argument");
  if (int rc = finish_rollout(p)) return rc;
  const CostParams C = make_cost(k_goal_dist, k_path_len, k_safe_dist, k_workspace, ws_limits);
  if (int rc = launch_evaluate(p, C)) return rc;
  static_assert(offsetof(HostOut, best) == sizeof(EvalResult) && offsetof(HostOut, real) == sizeof(EvalResult) + sizeof(DeviceBest) &&
                    offsetof(HostOut, real_path) == offsetof(HostOut, real) + sizeof(RealState),
                "HostOut members must be contiguous");
  if (p->world == 1 || p->p2p_ready) {
    if (int rc = wait_ticket(p, 0, p->ticket)) return rc;
    REQUIRE(!p->h_out->p2p_fail, PMAF_ERR_NCCL, "best-agent exchange: a peer's record never arrived");
  } else {
    if (int rc = d2h(p, &p->h_out->eval, p->eval.p, sizeof(EvalResult) + sizeof(DeviceBest))) return rc;
    CU(cudaStreamSynchronize(p->stream));
  }
  p->h_eval = p->h_out->eval, p->h_best = p->h_out->best;
  *best_index = p->h_eval.best_index;
  return 0;
}

// apply a pending obstacle feed on the device (consumers other than pmaf_tick's kernel)
static int flush_pending_feed(pmaf_planner *p) {
  if (p->pending_feed_n <= 0) return 0;
  const int n = p->pending_feed_n;
  p->pending_feed_n = 0;
  return launch(p, feed_kernel, dim3(1), dim3(256), 0, p->live_pos.p, (const double *)p->live_vel.p, n, p->pending_feed_freq);
}

extern "C" int pmaf_feed_obstacles(pmaf_planner *p, int n_feed, double frequency) {
  ENTER(p);
  NEED_INIT(p);
  const size_t n = p->h_live.size() / 7;  // entries of the list the device holds
  REQUIRE(n > 0, PMAF_ERR_STATE, "pmaf_feed_obstacles: no obstacle list has been passed yet (evaluate / move / reset / tick)");
  REQUIRE(n == (size_t)p->O, PMAF_ERR_STATE,
          "pmaf_feed_obstacles: the device-resident list has %zu entries, the planner %d (partial lists are re-uploaded)", n, p->O);
  REQUIRE(n_feed >= 0 && (size_t)n_feed <= n && frequency > 0.0, PMAF_ERR_ARG, "pmaf_feed_obstacles: n_feed=%d frequency=%g",
          n_feed, frequency);
  if (n_feed == 0) return 0;
  if (int rc = flush_pending_feed(p)) return rc;  // an earlier step nobody consumed yet
  // host mirror: the same IEEE operations as the device (no contraction on either side)
  double *pos = p->h_live.data();
  const double *vel = pos + 3 * n;
  for (int i = 0; i < 3 * n_feed; ++i) pos[i] += vel[i] / frequency;
  p->pending_feed_n = n_feed, p->pending_feed_freq = frequency;
  p->live_changed = true;
  return 0;
}

// upload a live obstacle list unless it is byte-identical to the last one uploaded
static int upload_live(pmaf_planner *p, int n_obs, const double *obs_pos, const double *obs_vel,
                       const double *obs_rad) {
  if (int rc = flush_pending_feed(p)) return rc;
  const size_t n = (size_t)n_obs;
  std::vector<double> &last = p->h_live;
  const bool same = p->upload_dedup && last.size() == 7 * n && memcmp(last.data(), obs_pos, 3 * n * sizeof(double)) == 0 &&
                    memcmp(last.data() + 3 * n, obs_vel, 3 * n * sizeof(double)) == 0 &&
                    memcmp(last.data() + 6 * n, obs_rad, n * sizeof(double)) == 0;
  if (same) return 0;
  p->live_changed = true;
  last.resize(7 * n);
  memcpy(last.data(), obs_pos, 3 * n * sizeof(double));
  memcpy(last.data() + 3 * n, obs_vel, 3 * n * sizeof(double));
  memcpy(last.data() + 6 * n, obs_rad, n * sizeof(double));
  if (int rc = stage_acquire(p, 7 * n)) return rc;
  memcpy(p->h_stage, last.data(), 7 * n * sizeof(double));
  if (n_obs == p->O) {  // the usual case: the device block has the staging layout, one copy
    if (int rc = h2d(p, p->live_pos.p, p->h_stage, 7 * n * sizeof(double))) return rc;
  } else {
    if (int rc = h2d(p, p->live_pos.p, p->h_stage, 3 * n * sizeof(double))) return rc;
    if (int rc = h2d(p, p->live_vel.p, p->h_stage + 3 * n, 3 * n * sizeof(double))) return rc;
    if (int rc = h2d(p, p->live_rad.p, p->h_stage + 6 * n, n * sizeof(double))) return rc;
  }
  return stage_release(p);
}

static RealArgs make_real_args(pmaf_planner *p, int n_obs, double delta_t, int steps, int agent_id_global);
static int launch_real(pmaf_planner *p, int n_obs, double delta_t, int steps, int agent_id_global) {
  PlannerDev d = make_dev(p);
  const RealArgs r = make_real_args(p, n_obs, delta_t, steps, agent_id_global);
  return launch(p, real_agent_kernel, dim3(1), dim3(32), 0, d, r);
}

extern "C" int pmaf_move_real_agent(pmaf_planner *p, int n_obs, const double *obs_pos, const double *obs_vel,
                                    const double *obs_rad, double delta_t, int steps, int agent_id) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(obs_pos && obs_vel && obs_rad, PMAF_ERR_ARG, "pmaf_move_real_agent: null obstacle array");
  REQUIRE(n_obs >= 1 && n_obs <= p->O, PMAF_ERR_ARG,
          "pmaf_move_real_agent: n_obs=%d but the planner was initialised with %d obstacles", n_obs, p->O);
  REQUIRE(agent_id >= 0 && agent_id < p->n_global, PMAF_ERR_ARG, "pmaf_move_real_agent: agent_id=%d out of range",
          agent_id);
  REQUIRE(p->h_best.present, PMAF_ERR_STATE,
          "pmaf_move_real_agent: no best agent yet (the reference dereferences a null best_agent_ here)");
  REQUIRE(steps >= 0, PMAF_ERR_ARG, "pmaf_move_real_agent: steps=%d", steps);
  if (int rc = upload_live(p, n_obs, obs_pos, obs_vel, obs_rad)) return rc;
  for (int done = 0; done < steps;) {
    const int chunk = std::min(steps - done, 256);
    if (int rc = launch_real(p, n_obs, delta_t, chunk, agent_id)) return rc;
    if (int rc = wait_ticket(p, 1, p->ticket)) return rc;
    p->h_real = p->h_out->real;
    p->real_path.insert(p->real_path.end(), p->h_out->real_path, p->h_out->real_path + (size_t)chunk * 3);
    done += chunk;
  }
  return 0;
}

static int refresh_obstacle_copy(pmaf_planner *p, int n_obs, const double *obs_pos, const double *obs_vel,
                                 const double *pos) {
  memcpy(p->h_obs_pos.data(), obs_pos, (size_t)n_obs * 3 * sizeof(double));
  memcpy(p->h_obs_vel.data(), obs_vel, (size_t)n_obs * 3 * sizeof(double));
  ObstacleImage im = layout_image(p->O, any_nonzero(p->h_obs_vel));
  im.nn_valid = wants_nn_table(p, im);
  if (im.bytes != p->img.bytes) CU(p->image.resize(im.bytes));
  p->img = im;
  refresh_margin_scale(p);
  update_margin(p, pos);
  p->image_current = false;
  return 0;
}

extern "C" int pmaf_reset_agents(pmaf_planner *p, const double pos[3], const double vel[3], int n_obs,
                                 const double *obs_pos, const double *obs_vel, const double *obs_rad) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(pos && vel && obs_pos && obs_vel && obs_rad, PMAF_ERR_ARG, "pmaf_reset_agents: null argument");
  REQUIRE(n_obs >= 0 && n_obs <= p->O, PMAF_ERR_ARG,
          "pmaf_reset_agents: n_obs=%d but the planner was initialised with %d obstacles (std::out_of_range in the "
          "reference)", n_obs, p->O);
  if (int rc = finish_rollout(p)) return rc;
  if (int rc = upload_live(p, n_obs, obs_pos, obs_vel, obs_rad)) return rc;
  if (int rc = refresh_obstacle_copy(p, n_obs, obs_pos, obs_vel, pos)) return rc;
  // the node passes getNextPosition() / getNextVelocity() (node:350): those are the real agent's state, which
  // the device already holds bit for bit — no upload then
  const bool from_real = memcmp(pos, p->h_real.pos, 3 * sizeof(double)) == 0 && memcmp(vel, p->h_real.vel, 3 * sizeof(double)) == 0;
  if (!from_real) {
    if (int rc = stage_acquire(p, 6)) return rc;
    memcpy(p->h_stage, pos, 3 * sizeof(double));
    memcpy(p->h_stage + 3, vel, 3 * sizeof(double));
    if (int rc = h2d(p, p->scratch.p, p->h_stage, 6 * sizeof(double))) return rc;
    if (int rc = stage_release(p)) return rc;
  }
  p->fused_valid = p->have_cost;
  if (int rc = launch_reset(p, true, from_real, p->scratch.p, n_obs, p->live_pos.p, p->live_vel.p, true, true))
    return rc;
  p->obstacles_advanced = false;
  p->agents_touched = true;
  return 0;
}

// RealArgs of one launch (shared by real_agent_kernel and tick_kernel)
static RealArgs make_real_args(pmaf_planner *p, int n_obs, double delta_t, int steps, int agent_id_global) {
  RealArgs r{};
  r.real = p->real.p, r.known = p->real_known.p, r.rot = p->real_rot.p, r.best = p->best.p;
  r.best_random = p->best_random.p;
  r.obs_pos = p->live_pos.p, r.obs_vel = p->live_vel.p, r.obs_rad = p->live_rad.p, r.n_obs = n_obs;
  r.delta_t = delta_t, r.steps = steps, r.agent_id = agent_id_global, r.eval = p->eval.p;
  r.path_out = p->real_path_out.p;
  for (int i = 0; i < 3; ++i) r.goal[i] = p->goal[i];
  r.host = p->h_out_dev, r.ticket = ++p->ticket;
  r.pub_eval = nullptr, r.pub_best = nullptr;
  p->ctr.d2h_bytes += sizeof(RealState) + (size_t)steps * 3 * sizeof(double);  // stored by the kernel into the host block
  return r;
}

extern "C" int pmaf_tick(pmaf_planner *p, const double *measured_pos, int n_obs, const double *obs_pos,
                         const double *obs_vel, const double *obs_rad, double delta_t, double k_goal_dist,
                         double k_path_len, double k_safe_dist, double k_workspace, const double ws_limits[6],
                         int *best_index, double next_pos[3], double next_vel[3]) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(ws_limits, PMAF_ERR_ARG, "pmaf_tick: null argument");
  REQUIRE(n_obs >= 1 && n_obs <= p->O, PMAF_ERR_ARG, "pmaf_tick: n_obs=%d (planner has %d obstacles)", n_obs, p->O);
  const bool device_list = !obs_pos && !obs_vel && !obs_rad;
  REQUIRE(device_list || (obs_pos && obs_vel && obs_rad), PMAF_ERR_ARG, "pmaf_tick: pass all three obstacle arrays or none");
  if (device_list) {  // the device-resident list, as the last upload / pmaf_feed_obstacles left it (host mirror: h_live)
    REQUIRE(p->h_live.size() == 7 * (size_t)n_obs, PMAF_ERR_STATE,
            "pmaf_tick: no device-resident obstacle list of %d entries (pass the arrays once first)", n_obs);
    obs_pos = p->h_live.data(), obs_vel = obs_pos + 3 * (size_t)n_obs, obs_rad = obs_pos + 6 * (size_t)n_obs;
  }
  if (measured_pos) {
    if (int rc = pmaf_set_real_position(p, measured_pos)) return rc;
  }
  // the previous rollout is ordered before everything below by the stream; no host wait for it
  if (int rc = harvest_timing(p, p->roll_slot, false)) return rc;
  // a pending obstacle feed rides along in the tick kernel when the caller's list is the fed one
  int feed_n = 0;
  if (p->pending_feed_n > 0 && (device_list || (p->upload_dedup && p->h_live.size() == 7 * (size_t)n_obs &&
                                                memcmp(p->h_live.data(), obs_pos, 3 * (size_t)n_obs * sizeof(double)) == 0)))
    feed_n = p->pending_feed_n, p->pending_feed_n = 0;
  if (!device_list) {
    if (int rc = upload_live(p, n_obs, obs_pos, obs_vel, obs_rad)) return rc;
  }
  const CostParams C = make_cost(k_goal_dist, k_path_len, k_safe_dist, k_workspace, ws_limits);
  // Two launches per tick: tick_kernel (evaluateAgents [+ best-agent exchange], moveRealEEAgent, obstacle image /
  // known flags for the reset) and the rollout, whose prologue is resetEEAgents.
  PlannerDev d = make_dev(p);
  if (!(p->fused_valid && p->have_cost && same_cost(C, p->last_cost))) {
    if (int rc = launch(p, workspace_cost_kernel, dim3((p->A + 127) / 128), dim3(128), 0, d, C)) return rc;
  }
  p->last_cost = C, p->have_cost = true;
  if (p->live_changed || !p->image_current) {
    if (int rc = refresh_obstacle_copy(p, n_obs, obs_pos, obs_vel, p->h_real.pos)) return rc;
  } else {
    update_margin(p, p->h_real.pos);
  }
  d = make_dev(p);  // the image layout may have changed
  TickArgs T{};
  T.feed_n = feed_n, T.feed_frequency = p->pending_feed_freq, T.feed_pos = p->live_pos.p, T.feed_vel = p->live_vel.p;
  T.best = p->best.p, T.best_random = p->best_random.p, T.rec = reinterpret_cast<ArgminRecord *>(p->rec.p);
  T.eval = p->eval.p, T.host = p->h_out_dev, T.world = p->world;
  T.p2p_status = (int *)&p->h_out_dev->p2p_fail;
  T.known_bits = p->known_bits.p, T.known_keep = p->known_keep.p;
  T.rebuild_image = p->image_current ? 0 : 1;
  T.rebuild_nn = (T.rebuild_image && p->img.nn_valid) ? 1 : 0;
  const bool copy_eval = p->world > 1 && !p->p2p_ready;  // the NCCL path's selection does not write the host block
  if (p->world == 1) {
    T.eval_mode = 0;
  } else if (p->p2p_ready) {
    REQUIRE(p->p2p_rank == p->rank && p->p2p_world == p->world, PMAF_ERR_STATE,
            "peer-memory exchange was set up for rank %d of %d, the shard is rank %d of %d", p->p2p_rank, p->p2p_world,
            p->rank, p->world);
    T.eval_mode = 1;
    for (int r = 0; r < p->world; ++r) T.xchg.peers[r] = p->peer_xchg[r];
    T.xchg.rank = p->rank, T.xchg.world = p->world, T.xchg.seq = ++p->xseq, T.xchg.stride = p2p_slot_stride();
    p->ctr.collectives++;
  } else {
    REQUIRE(p->nccl != nullptr, PMAF_ERR_STATE, "sharded planner without an exchange (pmaf_nccl_init or pmaf_p2p_import)");
    const int threads = p->A >= 1024 ? 1024 : std::max(32, ((p->A + 31) / 32) * 32);
    if (int rc = launch(p, evaluate_kernel, dim3(1), dim3(threads), 0, d, C, p->best.p, p->best_random.p, T.rec, p->eval.p,
                        0, (HostOut *)nullptr, 0ull))
      return rc;
    NC(g_nccl.AllGather(p->rec.p, p->rec_all.p, argmin_record_bytes(p->O), ncclChar, (ncclComm_t)p->nccl, p->stream));
    p->ctr.collectives++;
    T.eval_mode = 2, T.rec_all = p->rec_all.p;
  }
  T.eval_ticket = ++p->ticket;
  p->ctr.d2h_bytes += sizeof(EvalResult) + sizeof(DeviceBest);
  RealArgs R = make_real_args(p, n_obs, delta_t, 1, -1);
  if (!copy_eval) R.pub_eval = p->eval.p, R.pub_best = p->best.p;  // published with the real agent's state
  ResetArgs S{};
  S.real = p->real.p, S.from_real = 1, S.n_obs_update = n_obs, S.new_pos = p->live_pos.p, S.new_vel = p->live_vel.p;
  S.obs_pos = p->obs_pos.p, S.obs_vel = p->obs_vel.p, S.obs_rad = p->obs_rad.p, S.real_known = p->real_known.p;
  S.image = p->image.p, S.margin = p->margin;
  // one warp for the real agent, the rest for the costs / the image (at least one more warp)
  const int threads = std::min(256, std::max(64, ((p->A + 31) / 32) * 32 + 32));
  if (int rc = launch(p, tick_kernel, dim3(1), dim3(threads), 0, d, C, T, R, S)) return rc;
  p->image_current = true, p->live_changed = false;
  p->fused_valid = true;
  const unsigned long long real_ticket = p->ticket;  // the real agent publishes last
  if (copy_eval) {
    if (int rc = d2h(p, &p->h_out->eval, p->eval.p, sizeof(EvalResult) + sizeof(DeviceBest))) return rc;
    CU(cudaEventRecord(p->ev_d2h, p->stream));
  }
  p->obstacles_advanced = false;
  if (int rc = launch_rollout(p, true)) return rc;
  if (copy_eval) CU(cudaEventSynchronize(p->ev_d2h));
  if (int rc = wait_ticket(p, 1, real_ticket)) return rc;
  REQUIRE(!p->h_out->p2p_fail, PMAF_ERR_NCCL, "best-agent exchange: a peer's record never arrived");
  p->h_eval = p->h_out->eval, p->h_best = p->h_out->best, p->h_real = p->h_out->real;
  p->real_path.insert(p->real_path.end(), p->h_real.pos, p->h_real.pos + 3);
  if (best_index) *best_index = p->h_eval.best_index;
  for (int i = 0; i < 3; ++i) {
    if (next_pos) next_pos[i] = p->h_real.pos[i];
    if (next_vel) next_vel[i] = p->h_real.vel[i];
  }
  return 0;
}

// The in-process equivalent of launch/dry_run.launch (:9,41: the planner's `goals` output relayed back as its
// `position` input) in the reference's language: `ticks` planCallbacks (node:329-369) through the public
// entry points above, in the node's order, with HOST obstacle lists; between ticks the obstacle feed of
// dynamic_obstacle_node (:352-369) advances obstacles [0, n_feed) by vel / feed_frequency.
extern "C" int pmaf_dry_run(pmaf_planner *p, int ticks, int n_obs, double *obs_pos, const double *obs_vel,
                            const double *obs_rad, int n_feed, double feed_frequency, double delta_t, double k_goal_dist,
                            double k_path_len, double k_safe_dist, double k_workspace, const double ws_limits[6],
                            int flags, double *seconds, int *best, double *next_pos, double *next_vel) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(ticks >= 0 && obs_pos && obs_vel && obs_rad && ws_limits && n_feed >= 0 && n_feed <= n_obs, PMAF_ERR_ARG,
          "pmaf_dry_run: bad argument");
  double total = 0.0;
  REQUIRE(!(flags & (PMAF_DRY_RUN_PROFILE | PMAF_DRY_RUN_TICK_TIMES)) || seconds, PMAF_ERR_ARG,
          "pmaf_dry_run: PMAF_DRY_RUN_PROFILE needs seconds[7], PMAF_DRY_RUN_TICK_TIMES seconds[7 + ticks]");
  if (flags & PMAF_DRY_RUN_PROFILE)
    for (int k = 1; k < 7; ++k) seconds[k] = 0.0;
  for (int t = 0; t < ticks; ++t) {
    if (flags & PMAF_DRY_RUN_FLUSH_L2) {
      if (int rc = pmaf_flush_l2(p)) return rc;
    }
    if (flags & (PMAF_DRY_RUN_FLUSH_L2 | PMAF_DRY_RUN_WAIT_ROLLOUT)) {  // drain before the clock starts
      if (int rc = pmaf_stop_prediction(p)) return rc;
      CU(cudaStreamSynchronize(p->stream));  // the flush, too
    }
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    timespec tc = t0;
    auto lap = [&](int k) {  // PMAF_DRY_RUN_PROFILE: seconds[1 + k] accumulates the wall time of call k
      if (!(flags & PMAF_DRY_RUN_PROFILE)) return;
      timespec tn;
      clock_gettime(CLOCK_MONOTONIC, &tn);
      seconds[1 + k] += (double)(tn.tv_sec - tc.tv_sec) + 1e-9 * (double)(tn.tv_nsec - tc.tv_nsec);
      tc = tn;
    };
    int b = 0;
    double pos[3], vel[3];
    if (int rc = pmaf_stop_prediction(p)) return rc;
    lap(0);
    if (int rc = pmaf_evaluate_agents(p, n_obs, obs_pos, obs_vel, obs_rad, k_goal_dist, k_path_len, k_safe_dist, k_workspace,
                                      ws_limits, &b))
      return rc;
    lap(1);
    if (int rc = pmaf_move_real_agent(p, n_obs, obs_pos, obs_vel, obs_rad, delta_t, 1, b)) return rc;
    lap(2);
    if (int rc = pmaf_get_next_position(p, pos)) return rc;
    if (int rc = pmaf_get_next_velocity(p, vel)) return rc;
    if (int rc = pmaf_reset_agents(p, pos, vel, n_obs, obs_pos, obs_vel, obs_rad)) return rc;
    lap(3);
    if (int rc = pmaf_start_prediction(p)) return rc;
    lap(4);
    if (flags & PMAF_DRY_RUN_WAIT_ROLLOUT) {
      if (int rc = pmaf_stop_prediction(p)) return rc;
    }
    lap(5);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double tick_s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    total += tick_s;
    if (flags & PMAF_DRY_RUN_TICK_TIMES) seconds[7 + t] = tick_s;
    if (best) best[t] = b;
    for (int i = 0; i < 3; ++i) {
      if (next_pos) next_pos[3 * t + i] = pos[i];
      if (next_vel) next_vel[3 * t + i] = vel[i];
    }
    for (int i = 0; i < 3 * n_feed; ++i) obs_pos[i] += obs_vel[i] / feed_frequency;
    if ((flags & PMAF_DRY_RUN_DEVICE_FEED) && n_feed > 0) {  // the same step on the device-resident list
      if (int rc = pmaf_feed_obstacles(p, n_feed, feed_frequency)) return rc;
    }
  }
  if (seconds) *seconds = total;
  return 0;
}

// ---- getters ---------------------------------------------------------------------------------------------------------------
extern "C" int pmaf_get_num_agents(pmaf_planner *p, int *n) {
  ENTER(p);
  REQUIRE(n, PMAF_ERR_ARG, "null output");
  *n = p->A;
  return 0;
}
#define VEC3_GETTER(name, src)                          \
  extern "C" int name(pmaf_planner *p, double out[3]) { \
    ENTER(p);                                           \
    REQUIRE(out, PMAF_ERR_ARG, "null output");          \
    for (int i = 0; i < 3; ++i) out[i] = (src)[i];      \
    return 0;                                           \
  }
VEC3_GETTER(pmaf_get_next_position, p->h_real.pos)
VEC3_GETTER(pmaf_get_next_velocity, p->h_real.vel)
VEC3_GETTER(pmaf_get_ee_force, p->h_real.force)
VEC3_GETTER(pmaf_get_goal_position, p->goal)
VEC3_GETTER(pmaf_get_initial_position, p->mgr_init_pos)

extern "C" int pmaf_get_dist_from_goal(pmaf_planner *p, double *out) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  *out = norm3(sub3(ld3(p->goal), ld3(p->h_real.pos)));  // cf_manager.h:87-89 on the host mirror
  return 0;
}
extern "C" int pmaf_get_best_agent_type(pmaf_planner *p, int *out) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  *out = p->h_best.present ? p->h_best.type : -1;
  return 0;
}
extern "C" int pmaf_get_best_agent_id(pmaf_planner *p, int *out) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  *out = p->h_best.present ? p->h_best.id : 0;
  return 0;
}

static int fetch(pmaf_planner *p, void *dst, const void *src, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, p->stream));
  p->ctr.d2h_bytes += bytes;
  return 0;
}

extern "C" int pmaf_get_num_prediction_steps(pmaf_planner *p, int agent, int *out) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(out && agent >= 0 && agent < p->A, PMAF_ERR_ARG, "pmaf_get_num_prediction_steps: agent=%d", agent);
  if (int rc = finish_rollout(p)) return rc;
  if (int rc = fetch(p, out, p->n_path.p + agent, sizeof(int))) return rc;
  CU(cudaStreamSynchronize(p->stream));
  return 0;
}
extern "C" int pmaf_get_real_num_prediction_steps(pmaf_planner *p, int *out) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  *out = (int)(p->real_path.size() / 3);
  return 0;
}

extern "C" int pmaf_get_agent_summaries(pmaf_planner *p, int *steps, double *length, double *min_obs_dist,
                                        int *reached, double *pred_time_ns, int *agent_type) {
  ENTER(p);
  NEED_INIT(p);
  if (int rc = finish_rollout(p)) return rc;
  const size_t A = p->A;
  if (steps)
    if (int rc = fetch(p, steps, p->n_path.p, A * sizeof(int))) return rc;
  if (length)
    if (int rc = fetch(p, length, p->path_len.p, A * sizeof(double))) return rc;
  if (min_obs_dist)
    if (int rc = fetch(p, min_obs_dist, p->min_obs.p, A * sizeof(double))) return rc;
  if (reached)
    if (int rc = fetch(p, reached, p->reached.p, A * sizeof(int))) return rc;
  if (pred_time_ns)
    if (int rc = fetch(p, pred_time_ns, p->pred_time.p, A * sizeof(double))) return rc;
  CU(cudaStreamSynchronize(p->stream));
  if (agent_type)
    for (size_t a = 0; a < A; ++a) agent_type[a] = agent_type_of_index(p->first_agent + (int)a);
  return 0;
}

extern "C" int pmaf_get_predicted_paths(pmaf_planner *p, double *out, int stride) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(out && stride >= 1, PMAF_ERR_ARG, "pmaf_get_predicted_paths: bad argument");
  if (int rc = finish_rollout(p)) return rc;
  const size_t A = p->A, H = p->H;
  std::vector<int> n(A);
  std::vector<double> tmp(A * H * 3);
  if (int rc = fetch(p, n.data(), p->n_path.p, A * sizeof(int))) return rc;
  if (int rc = fetch(p, tmp.data(), p->paths.p, A * H * 3 * sizeof(double))) return rc;
  CU(cudaStreamSynchronize(p->stream));
  for (size_t a = 0; a < A; ++a) {
    const size_t rows = std::min<size_t>((size_t)n[a], (size_t)stride);
    memcpy(out + a * (size_t)stride * 3, tmp.data() + a * H * 3, rows * 3 * sizeof(double));
  }
  return 0;
}

extern "C" int pmaf_get_predicted_path(pmaf_planner *p, int agent, double *out, int max_points, int *n_points) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(out && n_points && agent >= 0 && agent < p->A && max_points >= 0, PMAF_ERR_ARG,
          "pmaf_get_predicted_path: bad argument");
  if (int rc = finish_rollout(p)) return rc;
  int n = 0;
  if (int rc = fetch(p, &n, p->n_path.p + agent, sizeof(int))) return rc;
  CU(cudaStreamSynchronize(p->stream));
  *n_points = n;
  const size_t rows = (size_t)std::min(n, max_points);
  if (rows) {
    if (int rc = fetch(p, out, p->paths.p + (size_t)agent * p->H * 3, rows * 3 * sizeof(double))) return rc;
    CU(cudaStreamSynchronize(p->stream));
  }
  return 0;
}

extern "C" int pmaf_get_agent_velocities(pmaf_planner *p, double *out) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  if (int rc = finish_rollout(p)) return rc;
  if (int rc = fetch(p, out, p->vel.p, (size_t)p->A * 3 * sizeof(double))) return rc;
  CU(cudaStreamSynchronize(p->stream));
  return 0;
}

extern "C" int pmaf_get_planned_trajectory(pmaf_planner *p, double *out, int max_points, int *n_points) {
  ENTER(p);
  REQUIRE(n_points, PMAF_ERR_ARG, "null output");
  const int n = (int)(p->real_path.size() / 3);
  *n_points = n;
  if (out && max_points > 0) memcpy(out, p->real_path.data(), (size_t)std::min(n, max_points) * 3 * sizeof(double));
  return 0;
}

extern "C" int pmaf_get_obstacle_state(pmaf_planner *p, int n_obs, int *known, double *rot) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(n_obs == p->O, PMAF_ERR_ARG, "pmaf_get_obstacle_state: n_obs=%d, planner has %d", n_obs, p->O);
  if (int rc = finish_rollout(p)) return rc;
  const size_t A = p->A, O = p->O, KW = p->known_words;
  if (known) {
    std::vector<uint32_t> w(A * KW);
    std::vector<unsigned char> rk(O);
    if (int rc = fetch(p, w.data(), p->known.p, A * KW * sizeof(uint32_t))) return rc;
    if (int rc = fetch(p, rk.data(), p->real_known.p, O)) return rc;
    CU(cudaStreamSynchronize(p->stream));
    for (size_t a = 0; a < A; ++a)
      for (size_t i = 0; i < O; ++i) known[a * O + i] = (w[a * KW + (i >> 5)] >> (i & 31)) & 1u;
    for (size_t i = 0; i < O; ++i) known[A * O + i] = rk[i] != 0;
  }
  if (rot) {
    if (int rc = fetch(p, rot, p->rot.p, A * O * 3 * sizeof(double))) return rc;
    if (int rc = fetch(p, rot + A * O * 3, p->real_rot.p, O * 3 * sizeof(double))) return rc;
    CU(cudaStreamSynchronize(p->stream));
  }
  return 0;
}

extern "C" int pmaf_get_costs(pmaf_planner *p, double *costs) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(costs, PMAF_ERR_ARG, "null output");
  if (int rc = fetch(p, costs, p->cost.p, (size_t)p->A * sizeof(double))) return rc;
  CU(cudaStreamSynchronize(p->stream));
  return 0;
}

extern "C" int pmaf_get_counters(pmaf_planner *p, pmaf_counters *out) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  if (p->initialized) {
    if (int rc = finish_rollout(p)) return rc;
    if (int rc = fetch(p, p->h_out->steps, p->step_counter.p, 16 * sizeof(unsigned long long))) return rc;
    CU(cudaStreamSynchronize(p->stream));
    p->ctr.agent_steps = p->h_out->steps[0];
    p->ctr.agent_steps_total = p->h_out->steps[1];
    p->ctr.general_steps_total = p->h_out->steps[2];
  }
  *out = p->ctr;
  return 0;
}

// Developer statistics (libraries built with -DPMAF_FAST_STATS): how often each reason sent a step of the
// latency build to the general step; out[12], all zero otherwise. Call after pmaf_get_counters.
extern "C" int pmaf_get_fast_stats(pmaf_planner *p, uint64_t out[12]) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  for (int i = 0; i < 12; ++i) out[i] = p->h_out->steps[4 + i];
  return 0;
}

extern "C" int pmaf_set_tuning(pmaf_planner *p, int lanes_per_agent, int block_threads, int occupancy) {
  ENTER(p);
  REQUIRE(lanes_per_agent == 0 || lanes_per_agent == 4 || lanes_per_agent == 8 || lanes_per_agent == 16 ||
              lanes_per_agent == 32,
          PMAF_ERR_ARG, "pmaf_set_tuning: lanes_per_agent must be 0, 4, 8, 16 or 32");
  REQUIRE(block_threads == 0 || (block_threads >= 32 && block_threads <= 256 && block_threads % 32 == 0),
          PMAF_ERR_ARG, "pmaf_set_tuning: block_threads must be 0 or a multiple of 32 in [32, 256]");
  REQUIRE(occupancy == 0 || occupancy == 1 || occupancy == 3 || occupancy == 4, PMAF_ERR_ARG,
          "pmaf_set_tuning: occupancy must be 0 (auto), 1, 3 or 4");
  p->tune_lpa = lanes_per_agent, p->tune_block = block_threads, p->tune_occ = occupancy;
  return 0;
}

extern "C" int pmaf_set_rollout_timing(pmaf_planner *p, int on) {
  ENTER(p);
  if (int rc = finish_rollout(p)) return rc;
  p->time_rollouts = on != 0;
  return 0;
}

extern "C" int pmaf_set_upload_dedup(pmaf_planner *p, int dedup) {
  ENTER(p);
  p->upload_dedup = dedup != 0;
  return 0;
}

extern "C" int pmaf_timer_start(pmaf_planner *p) {
  ENTER(p);
  CU(cudaEventRecord(p->ev_t0, p->stream));
  return 0;
}

extern "C" int pmaf_timer_stop(pmaf_planner *p, double *elapsed_ms) {
  ENTER(p);
  REQUIRE(elapsed_ms, PMAF_ERR_ARG, "null output");
  CU(cudaEventRecord(p->ev_t1, p->stream));
  CU(cudaEventSynchronize(p->ev_t1));
  if (p->rollout_pending) {
    p->rollout_pending = false;  // the stream is idle now
    if (int rc = harvest_timing(p, p->roll_slot ^ 1, true)) return rc;
    if (int rc = harvest_timing(p, p->roll_slot, true)) return rc;
  }
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, p->ev_t0, p->ev_t1));
  *elapsed_ms = ms;
  return 0;
}

extern "C" int pmaf_flush_l2(pmaf_planner *p) {
  ENTER(p);
  const size_t bytes = (size_t)256 << 20;
  CU(p->l2_scratch.resize(bytes));
  CU(cudaMemsetAsync(p->l2_scratch.p, 0x5a, bytes, p->stream));
  return 0;
}

namespace pmaf {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3,
         a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
    a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
}  // namespace pmaf

extern "C" int pmaf_measure_fp64_peak(pmaf_planner *p, double *tflops) {
  ENTER(p);
  REQUIRE(tflops, PMAF_ERR_ARG, "null output");
  const int blocks = 148 * 8, threads = 256, iters = 1 << 15;
  DevBuf<double> out;
  CU(out.resize((size_t)blocks * threads));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(p->ev_t0, p->stream));
    if (int rc = launch(p, fp64_peak_kernel, dim3(blocks), dim3(threads), 0, out.p, iters)) return rc;
    CU(cudaEventRecord(p->ev_t1, p->stream));
    CU(cudaEventSynchronize(p->ev_t1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, p->ev_t0, p->ev_t1));
    if (rep > 0) best_ms = std::min(best_ms, ms);
  }
  out.release();
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best_ms * 1e-3) / 1e12;
  return 0;
}

// ---- FastMath self-test ---------------------------------------------------------------------------------------
namespace pmaf {
__device__ __forceinline__ unsigned long long xorshift(unsigned long long &s) {
  s ^= s << 13, s ^= s >> 7, s ^= s << 17;
  return s;
}
__device__ __forceinline__ double random_double(unsigned long long &s, int emin, int emax, bool allow_negative) {
  const unsigned long long mant = xorshift(s) & ((1ull << 52) - 1);
  const int e = emin + (int)(xorshift(s) % (unsigned long long)(emax - emin + 1));
  const unsigned long long sign = allow_negative ? (xorshift(s) & 1ull) << 63 : 0ull;
  return __longlong_as_double((long long)(sign | ((unsigned long long)(e + 1023) << 52) | mant));
}
// out[0] sqrt mismatches, [1] div mismatches, [2] div3 mismatches, [3] samples with the range flag raised,
// [4] samples compared
__global__ void math_selftest_kernel(unsigned long long seed, int iters, unsigned long long *out) {
  unsigned long long s = seed * 0x9E3779B97F4A7C15ull + (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0xD1B54A32D192ED03ull + 1;
  for (int w = 0; w < 8; ++w) xorshift(s);
  unsigned long long bad_sqrt = 0, bad_div = 0, bad_div3 = 0, flagged = 0, compared = 0;
  for (int it = 0; it < iters; ++it) {
    const int kind = it & 7;
    double a, b;
    if (kind == 0) {  // generic
      a = random_double(s, -40, 40, true), b = random_double(s, -40, 40, false);
    } else if (kind == 1) {  // quotient next to a rounding boundary: a = RN(b * (q + ulp/2)) +- ulp
      b = random_double(s, -6, 6, false);
      const double q = random_double(s, -6, 6, false);
      const double half_ulp = __longlong_as_double(__double_as_longlong(q) + 1) - q;
      a = fma(b, 0.5 * half_ulp, b * q);
      const long long adj = (long long)(xorshift(s) % 3) - 1;
      a = __longlong_as_double(__double_as_longlong(a) + adj);
    } else if (kind == 2) {  // divisor significand all ones / near a power of two
      const unsigned long long mant = (xorshift(s) & 1) ? ((1ull << 52) - 1 - (xorshift(s) & 15)) : (xorshift(s) & 15);
      b = __longlong_as_double((long long)(((unsigned long long)(1023 + (int)(xorshift(s) % 21) - 10) << 52) | mant));
      a = random_double(s, -10, 10, true);
    } else if (kind == 3) {  // component of a vector over its norm: |a| <= b
      b = random_double(s, -14, 4, false);
      a = b * ((double)(xorshift(s) >> 11) * (1.0 / 9007199254740992.0)) * ((xorshift(s) & 1) ? 1.0 : -1.0);
    } else if (kind == 4) {  // square root of an exact square and its neighbours
      const double r = random_double(s, -20, 20, false);
      const double r26 = __longlong_as_double(__double_as_longlong(r) & ~((1ll << 27) - 1));  // 26-bit r: r*r exact
      a = r26 * r26;
      a = __longlong_as_double(__double_as_longlong(a) + (long long)(xorshift(s) % 3) - 1);
      b = random_double(s, -3, 3, false);
    } else if (kind == 5) {  // squared norms of metre-scale vectors
      const double x = random_double(s, -12, 2, true), y = random_double(s, -12, 2, true), z = random_double(s, -12, 2, true);
      a = (x * x + y * y) + z * z;
      b = sqrt(a);
    } else if (kind == 6) {  // operands that must raise the flag
      const int pick = (int)(xorshift(s) % 6);
      a = pick == 0 ? 0.0 : pick == 1 ? -1.0 : pick == 2 ? 1e-300 : pick == 3 ? 1e300 : pick == 4 ? __longlong_as_double(0x7ff8000000000000ll) : random_double(s, -2, 2, true);
      b = pick == 5 ? 0.0 : random_double(s, -2, 2, false);
    } else {  // wide exponents inside the proven range
      a = random_double(s, -590, 590, true), b = random_double(s, -295, 295, false);
    }
    const double x = fabs(a);
    {
      FastMath fm;
      const double got = fm.sqrt_(x);
      if (fm.bad()) ++flagged;
      else if (__double_as_longlong(got) != __double_as_longlong(sqrt(x))) ++bad_sqrt;
    }
    {
      FastMath fm;
      const double got = fm.div_(a, b);
      if (fm.bad()) ++flagged;
      else if (__double_as_longlong(got) != __double_as_longlong(a / b)) ++bad_div;
    }
    {
      FastMath fm;
      const v3 num = mk3(a, 0.37 * a, -a * 1.9);
      const v3 got = fm.div3_(num, b);
      if (fm.bad()) ++flagged;
      else if (__double_as_longlong(got.x) != __double_as_longlong(num.x / b) ||
               __double_as_longlong(got.y) != __double_as_longlong(num.y / b) ||
               __double_as_longlong(got.z) != __double_as_longlong(num.z / b))
        ++bad_div3;
    }
    {  // fused root + reciprocal of the root: a normalisation (components bounded by the norm) and a free numerator
      FastMath fm;
      double s_, y_;
      fm.sqrt_rcp_(x, s_, y_);
      const double frac = (double)(xorshift(s) >> 11) * (1.0 / 9007199254740992.0);
      const v3 num = mk3(frac * s_, -s_, (kind & 1) ? 0.0 : s_ * 0x1p-40 * frac);
      const v3 got = fm.quot3_(num, s_, y_);
      const double gq = fm.quot_(a, s_, y_);
      if (fm.bad()) ++flagged;
      else {
        if (__double_as_longlong(s_) != __double_as_longlong(sqrt(x))) ++bad_sqrt;
        if (__double_as_longlong(gq) != __double_as_longlong(a / s_)) ++bad_div;
        if (__double_as_longlong(got.x) != __double_as_longlong(num.x / s_) ||
            __double_as_longlong(got.y) != __double_as_longlong(num.y / s_) ||
            __double_as_longlong(got.z) != __double_as_longlong(num.z / s_))
          ++bad_div3;
      }
    }
    compared += 4;
  }
  atomicAdd(out + 0, bad_sqrt), atomicAdd(out + 1, bad_div), atomicAdd(out + 2, bad_div3);
  atomicAdd(out + 3, flagged), atomicAdd(out + 4, compared);
}
}  // namespace pmaf

extern "C" int pmaf_selftest_math(pmaf_planner *p, uint64_t samples, uint64_t seed, uint64_t out[5]) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  DevBuf<unsigned long long> d;
  CU(d.resize(5));
  CU(cudaMemsetAsync(d.p, 0, 5 * sizeof(unsigned long long), p->stream));
  const int blocks = 148 * 8, threads = 256;
  const int iters = (int)std::max<uint64_t>(8, samples / ((uint64_t)blocks * threads));
  if (int rc = launch(p, math_selftest_kernel, dim3(blocks), dim3(threads), 0, (unsigned long long)seed, iters, d.p)) return rc;
  unsigned long long h[5];
  CU(cudaMemcpyAsync(h, d.p, sizeof h, cudaMemcpyDeviceToHost, p->stream));
  CU(cudaStreamSynchronize(p->stream));
  for (int i = 0; i < 5; ++i) out[i] = h[i];
  d.release();
  return 0;
}

// Per-section cycle counters of the first 64 agents' last rollout (all zero unless the library was
// built with -DPMAF_SECTION_TIMERS); out[64][12].
extern "C" int pmaf_get_section_cycles(pmaf_planner *p, long long *out) {
  ENTER(p);
  REQUIRE(out, PMAF_ERR_ARG, "null output");
  if (!p->section_cycles.p) {
    memset(out, 0, 64 * 12 * sizeof(long long));
    return 0;
  }
  if (int rc = finish_rollout(p)) return rc;
  CU(cudaMemcpyAsync(out, p->section_cycles.p, 64 * 12 * sizeof(long long), cudaMemcpyDeviceToHost, p->stream));
  CU(cudaStreamSynchronize(p->stream));
  return 0;
}

extern "C" int pmaf_get_best_paths(pmaf_planner *p, int k, int stride, int max_points, int *agent_index, int *n_points,
                                   double *paths) {
  ENTER(p);
  NEED_INIT(p);
  REQUIRE(k >= 1 && k <= 64 && stride >= 1 && max_points >= 1 && agent_index && n_points && paths, PMAF_ERR_ARG,
          "pmaf_get_best_paths: bad argument (1 <= k <= 64, stride >= 1)");
  REQUIRE(p->have_cost, PMAF_ERR_STATE, "pmaf_get_best_paths: no evaluate_agents yet");
  if (int rc = finish_rollout(p)) return rc;
  DevBuf<int> d_idx;
  CU(d_idx.resize(k));
  PlannerDev d = make_dev(p);
  const int threads = p->A >= 1024 ? 1024 : std::max(32, ((p->A + 31) / 32) * 32);
  if (int rc = launch(p, topk_kernel, dim3(1), dim3(threads), 0, d, k, d_idx.p)) return rc;
  if (int rc = fetch(p, agent_index, d_idx.p, (size_t)k * sizeof(int))) return rc;
  CU(cudaStreamSynchronize(p->stream));
  d_idx.release();
  std::vector<double> row((size_t)p->H * 3);
  for (int r = 0; r < k; ++r) {
    n_points[r] = 0;
    const int a = agent_index[r];
    if (a < 0) continue;
    int n = 0;
    if (int rc = fetch(p, &n, p->n_path.p + a, sizeof(int))) return rc;
    CU(cudaStreamSynchronize(p->stream));
    if (int rc = fetch(p, row.data(), p->paths.p + (size_t)a * p->H * 3, (size_t)n * 3 * sizeof(double))) return rc;
    CU(cudaStreamSynchronize(p->stream));
    int m = 0;  // every stride-th point, always including the last one
    for (int q = 0; q < n && m < max_points; q += stride, ++m) memcpy(paths + ((size_t)r * max_points + m) * 3, &row[3 * (size_t)q], 24);
    if (n > 0 && (n - 1) % stride != 0 && m < max_points) memcpy(paths + ((size_t)r * max_points + m++) * 3, &row[3 * (size_t)(n - 1)], 24);
    n_points[r] = m;
    agent_index[r] = a + p->first_agent;
  }
  return 0;
}

// ---- downstream kinematics -----------------------------------------------------------------------
extern "C" void pmaf_panda_joint_limits(double q_lo[7], double q_hi[7]) {
  // src/costp_controller.cpp:41-44
  const double lo[7] = {-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
  const double hi[7] = {2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
  for (int i = 0; i < 7; ++i) q_lo[i] = lo[i], q_hi[i] = hi[i];
}

extern "C" int pmaf_dq_kinematics(pmaf_planner *p, const double base_dq[8], const double q[7], double pose[8],
                                  double pose_jacobian[56], double geom_jacobian[42]) {
  ENTER(p);
  REQUIRE(base_dq && q && pose && pose_jacobian && geom_jacobian, PMAF_ERR_ARG, "pmaf_dq_kinematics: null argument");
  DevBuf<double> d;
  CU(d.resize(8 + 7 + 8 + 56 + 42));
  double h[15];
  memcpy(h, base_dq, 8 * sizeof(double)), memcpy(h + 8, q, 7 * sizeof(double));
  CU(cudaMemcpyAsync(d.p, h, sizeof h, cudaMemcpyHostToDevice, p->stream));
  if (int rc = launch(p, dq_probe_kernel, dim3(1), dim3(32), 0, (const double *)d.p, (const double *)(d.p + 8), d.p + 15,
                      d.p + 23, d.p + 79))
    return rc;
  double out[8 + 56 + 42];
  CU(cudaMemcpyAsync(out, d.p + 15, sizeof out, cudaMemcpyDeviceToHost, p->stream));
  CU(cudaStreamSynchronize(p->stream));
  memcpy(pose, out, 8 * sizeof(double)), memcpy(pose_jacobian, out + 8, 56 * sizeof(double));
  memcpy(geom_jacobian, out + 64, 42 * sizeof(double));
  d.release();
  return 0;
}

extern "C" int pmaf_score_paths(pmaf_planner *p, int k, const double base_dq[8], const double q_start[7], const double q_lo[7],
                                const double q_hi[7], double damping, double tol_pos, int *agent_index, pmaf_path_score *out) {
  ENTER(p);
  NEED_INIT(p);
  static_assert(sizeof(pmaf_path_score) == sizeof(PathScore), "pmaf_path_score mirrors PathScore");
  REQUIRE(base_dq && q_start && q_lo && q_hi && out && k <= 64 && damping > 0.0, PMAF_ERR_ARG,
          "pmaf_score_paths: bad argument (k <= 64, damping > 0)");
  REQUIRE(k <= 0 || agent_index, PMAF_ERR_ARG, "pmaf_score_paths: agent_index is required for k >= 1");
  REQUIRE(k <= 0 || p->have_cost, PMAF_ERR_STATE, "pmaf_score_paths: no evaluate_agents yet");
  if (int rc = finish_rollout(p)) return rc;
  const int n = k > 0 ? k : p->A;
  DevBuf<int> d_idx;
  DevBuf<PathScore> d_out;
  CU(d_out.resize(n));
  PlannerDev d = make_dev(p);
  if (k > 0) {  // the k cheapest agents of the last evaluate (local indices, -1 padded)
    CU(d_idx.resize(k));
    const int threads = p->A >= 1024 ? 1024 : std::max(32, ((p->A + 31) / 32) * 32);
    if (int rc = launch(p, topk_kernel, dim3(1), dim3(threads), 0, d, k, d_idx.p)) return rc;
  }
  ScoreArgs S{};
  S.paths = p->paths.p, S.n_path = p->n_path.p, S.max_steps = p->H, S.agent_index = k > 0 ? d_idx.p : nullptr, S.n = n;
  for (int i = 0; i < 8; ++i) S.base[i] = base_dq[i];
  for (int i = 0; i < 7; ++i) S.q_start[i] = q_start[i], S.q_lo[i] = q_lo[i], S.q_hi[i] = q_hi[i];
  S.damping = damping, S.tol_pos = tol_pos, S.out = d_out.p;
  if (int rc = launch(p, dq_score_kernel, dim3((n + 127) / 128), dim3(128), 0, S)) return rc;
  if (int rc = fetch(p, out, d_out.p, (size_t)n * sizeof(PathScore))) return rc;
  if (k > 0) {
    if (int rc = fetch(p, agent_index, d_idx.p, (size_t)k * sizeof(int))) return rc;
  }
  CU(cudaStreamSynchronize(p->stream));
  if (k > 0)
    for (int r = 0; r < k; ++r)
      if (agent_index[r] >= 0) agent_index[r] += p->first_agent;
  if (k <= 0 && agent_index)
    for (int a = 0; a < n; ++a) agent_index[a] = p->first_agent + a;
  d_idx.release(), d_out.release();
  return 0;
}
