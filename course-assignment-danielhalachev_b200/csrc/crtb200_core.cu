// crtb200_core.cu -- C ABI (include/crtb200.h) of the B200-native renderer core: host flattener H1 (reference trees ->
// stack-free visiting-order layout), device memory management, and the per-frame launch sequence of the wavefront
// kernels in crt_kernels.cuh.  No CPU rendering path exists in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/crtb200.h"
#include "crt_kernels.cuh"

using namespace crtd;

#ifndef CRT_REFILL
#define CRT_REFILL 16  // refill a warp when at least this many lanes are idle (8 until run r2aj: hw14 2.92 -> 2.77 ms)
#endif
#ifndef CRT_REFILL_CLOSEST
#define CRT_REFILL_CLOSEST CRT_REFILL
#endif
#ifndef CRT_REFILL_SHADOW
#define CRT_REFILL_SHADOW CRT_REFILL
#endif
#ifndef CRT_LOOP_MODE
#define CRT_LOOP_MODE 2  // 0 = while-while, 1 = merged loop, 2 = node phase + warp-cooperative triangle phase (crt_kernels.cuh)
#endif

struct crtb200_ctx;
static thread_local std::string g_error;
static thread_local crtb200_ctx *g_error_ctx = nullptr;  // context of the ABI call running on this thread (ErrScope)
static void note_ctx_error(crtb200_ctx *c, const std::string &msg);
static int fail(int code, const std::string &msg) {
  g_error = msg;
  if (g_error_ctx) note_ctx_error(g_error_ctx, msg);
  return code;
}
struct ErrScope {
  crtb200_ctx *prev;
  explicit ErrScope(crtb200_ctx *c) : prev(g_error_ctx) { g_error_ctx = c; }
  ~ErrScope() { g_error_ctx = prev; }
};

// One helper thread per peer GPU of a multi-GPU context: launches are asynchronous, but issuing a frame's ~20 launches
// takes the host ~0.1 ms per GPU, which eight GPUs cannot afford serially when a shard renders in ~0.5 ms.
struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<void()> job;
  bool has = false, quit = false;
  Worker() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return has || quit; });
        if (quit) return;
        lk.unlock();
        job();
        lk.lock();
        has = false;
        cv.notify_all();
      }
    });
  }
  void post(std::function<void()> f) {
    std::unique_lock<std::mutex> lk(m);
    job = std::move(f);
    has = true;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return !has; });
  }
  ~Worker() {
    {
      std::unique_lock<std::mutex> lk(m);
      quit = true;
      cv.notify_all();
    }
    th.join();
  }
};
#define CUDA_TRY(expr)                                                                                       \
  do {                                                                                                       \
    cudaError_t e_ = (expr);                                                                                 \
    if (e_ != cudaSuccess)                                                                                   \
      return fail(CRTB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorName(e_) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T> &v) {
    cudaError_t e = ensure(std::max<size_t>(v.size(), 1));
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct crtb200_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> kev;  // per-launch event pairs: K2 / K3 timing
  std::vector<int> kev_kind;     // 0 = closest, 1 = shadow (one entry per pair)
  size_t kev_used = 0;
  bool have_scene = false;
  uint64_t queue_budget = 16ull << 30;

  // scene
  DScene sc{};
  // nodes / leaf_refs / tri_geom (everything the traversal kernels chase) live in ONE allocation so that a single L2
  // access-policy window can mark them persisting: the ~1 GB of ray / hit / colour queues a 4K frame streams through
  // the 126 MB L2 then stops evicting the scene between and during the traversal kernels
  DevBuf<uint8_t> arena;
  float4 *nodes_p = nullptr, *tri_geom_p = nullptr;
  uint32_t *leaf_refs_p = nullptr;
  size_t arena_used = 0;
  size_t l2_persist_max = 0, l2_window_max = 0;
  int l2_persist = 0;  // CRT_L2_PERSIST: 0 off (default, measured best), 1 arena persisting / rest of it streaming, 2 arena persisting / normal, 3 nodes only
  size_t nodes_bytes = 0;
  DevBuf<float4> vtx_normal;
  bool nested_ok = false;    // every child box of the uploaded mesh trees lies inside its parent's and no tree is deeper
                             // than the k_coop LIFO allows: the order-free walk of k_coop and the subtree culling are exact
  int tail_iters = 16;       // tail hand-off (crt_kernels.cuh): once a traversal kernel's queue is dry, walks longer than a
                             // falling threshold (512 node-phase iterations, halved every 4 rounds, never below this
                             // floor) go to k_coop.  CRT_TAIL_ITERS overrides (tools / tests): 0 = every walk still
                             // running (the threshold still falls from 512), -1 = off
  int tail_start = 512;      // CRT_TAIL_START (tests): the threshold's starting value, shadow pass
  int tail_start_closest = 256;  // CRT_TAIL_START_CLOSEST: ... closest-hit launches (runs r2aq, r2at: 10M 2.02 -> 1.95 ms against 512, hw07 0.96 -> 1.00; 128: 1.93 / 1.08)
  int tail_small = 32768;    // CRT_TAIL_SMALL: launches of at most this many rays hand off at the floor from the start
  int tail_cap = 8;          // hand-off capacity per launch, in walks per resident k_coop warp (CRT_TAIL_CAP; tests use
                             // a huge value so that every walk goes through k_coop)
  DevBuf<uint32_t> top_refs;
  DevBuf<uint4> tri_shade;
  DevBuf<float2> vtx_uv;
  DevBuf<DMesh> meshes;
  DevBuf<DMaterial> materials;
  DevBuf<DTexture> textures;
  DevBuf<float> texels;
  DevBuf<DLight> lights;
  bool has_reflective = false, has_refractive = false;
  uint64_t scene_bytes = 0;

  // frame
  DevBuf<float> frame;   // persistent colour buffer (RayTracer::colorBuffer, RayTracer.h:69)
  DevBuf<uint8_t> frame8;
  // crtb200_render_frames: frames alternate between (frame, frame8) and (frame_b, frame8_b); the device->host copy of
  // frame f runs on copy_stream while frame f + 1 is rendered into the other pair
  DevBuf<float> frame_b;
  DevBuf<uint8_t> frame8_b;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  DevBuf<HitRec> hits;
  DevBuf<uint8_t> mask;
  std::vector<crtb200_rect> mask_rects;
  bool mask_valid = false, mask_needed = false;

  // queues: `concurrency` independent buffer sets, each with its own stream.  Consecutive chunks of a frame go to the
  // sets round-robin, so the latency-bound tails of one chunk's persistent kernels (and the small secondary-level
  // launches of reflective / refractive scenes) are filled by the other chunks' kernels.
  struct QueueSet {
    Levels lv{};
    DevBuf<float4> ray_o, ray_d, color, dq;
    DevBuf<uint32_t> hit_tri, ctl;  // ctl = counts | work cursors | hand-off counters, zeroed by one memset per chunk
    DevBuf<float> hit_t;
    DevBuf<uint4> comb;
    DevBuf<uint8_t> vis;
    DevBuf<uint4> ovf;          // walks handed off to k_coop (3 x uint4 per record)
    uint32_t *work = nullptr;   // into ctl
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // early k_coop pass: low-priority side stream, forked before and joined behind every traversal launch
    cudaStream_t coop_stream = nullptr;
    cudaEvent_t coop_fork = nullptr, coop_join = nullptr;
    void release() {
      ray_o.release(); ray_d.release(); color.release(); dq.release(); hit_tri.release(); ctl.release();
      hit_t.release(); comb.release(); vis.release(); ovf.release();
    }
  };
  std::vector<QueueSet> sets;
  uint32_t concurrency = 2;  // chunk streams of a host-bound frame (tools/e2e_time.py, run r2aq: 2 streams with one chunk each
                             // at 1080p, two each at 4K; 4 streams x 2 chunks cost 1080p frames 15-17 %)
  cudaEvent_t fork_ev = nullptr;
  DevBuf<unsigned long long> stats_dev;
  uint32_t cap_items = 0;
  uint32_t cap_depth = 0xFFFFFFFFu;
  uint32_t cap_sets = 0;
  // blocks per SM of k_coop's early pass (0 = final pass only); env CRT_COOP_EARLY.  Used for the primary level and the
  // shadow pass of frames that run on ONE stream (device frames, tile shards): with several chunk streams the tails
  // already overlap other chunks' work and waiting blocks would only take SM resources from them (run r2ad: 4.1 -> 8 ms);
  // the small secondary levels pay more for the extra launches and joins than they gain (hw11_room +2 %)
  int coop_early = CRT_COOP_MIN_BLOCKS;
  uint32_t n_chunks = 0;  // chunks of the current plan (chunk k runs on set k % cap_sets)
  cudaStream_t band_stream = nullptr;  // device -> host band copies of a chunked host-bound frame
  cudaEvent_t band_done = nullptr;
  std::vector<cudaEvent_t> chunk_done;
  std::vector<cudaEvent_t> dbg_ev;  // CRT_CHUNK_TIMES (tools): per chunk, end of k_store and end of its band copy
  uint32_t dbg_chunks = 0;

  int blocks_closest = 0, blocks_closest_sec = 0, blocks_shadow = 0, blocks_coop = 0;
  crtb200_stats last{};
  bool last_pending = false;

  // multi-GPU (crtb200_create_multi): this context is the primary (device_ids[0]).  A frame's 8x4-pixel tiles are dealt
  // round-robin over primary + peers (scene replicated); every GPU's k_store writes its pixels STRAIGHT into the
  // primary's frame buffers through peer-mapped pointers (NVLink / NVSwitch), so no slab, gather or assemble pass exists.
  std::vector<crtb200_ctx *> peers;
  std::vector<std::unique_ptr<Worker>> workers;
  std::vector<int> worker_rc;
  std::vector<std::string> worker_msg;
  cudaEvent_t multi_fork = nullptr;  // primary stream position the peers must wait for before touching the frame
  cudaEvent_t multi_done = nullptr;  // (on a peer) its shard is stored
  std::string error;                 // last error of a call on this context (crtb200_last_error_ctx)
};
static void note_ctx_error(crtb200_ctx *c, const std::string &msg) { c->error = msg; }

extern "C" {

uint32_t crtb200_abi_version(void) { return CRTB200_ABI_VERSION; }
const char *crtb200_last_error(void) { return g_error.c_str(); }

int crtb200_device_count(int *count) {
  if (!count) return fail(CRTB200_ERR_ARG, "count is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(CRTB200_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = n;
  return CRTB200_OK;
}

int crtb200_create(int device, crtb200_ctx **out) {
  if (!out) return fail(CRTB200_ERR_ARG, "out is null");
  *out = nullptr;
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(CRTB200_ERR_CUDA, "no such CUDA device (this library has no CPU fallback)");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(CRTB200_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100-class; libcrtb200 is built for sm_100a only");
  crtb200_ctx *c = new crtb200_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return fail(CRTB200_ERR_CUDA, "cudaStreamCreate failed");
  }
  for (auto &e : c->ev) cudaEventCreate(&e);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_closest<true, false, CRT_REFILL_CLOSEST, CRT_LOOP_MODE, false, false>, CRT_TRAV_BLOCK, 0);
  c->blocks_closest = std::max(1, occ) * c->sm_count;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_closest<false, false, CRT_REFILL_CLOSEST, CRT_LOOP_MODE, false, false>, CRT_TRAV_BLOCK, 0);
  c->blocks_closest_sec = std::max(1, occ) * c->sm_count;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shadow<0, CRT_REFILL_SHADOW, CRT_LOOP_MODE, false, false>, CRT_TRAV_BLOCK, 0);
  c->blocks_shadow = std::max(1, occ) * c->sm_count;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_coop<true, false, true, CRT_COOP_GROUP>, 32 * CRT_COOP_WARPS, 0);
  c->blocks_coop = std::max(1, occ) * c->sm_count;
  if (const char *env = getenv("CRT_BLOCKS_PER_SM")) {  // tuning only (tools/): resident persistent CTAs per SM
    const int b = atoi(env);
    if (b > 0) c->blocks_closest = c->blocks_closest_sec = c->blocks_shadow = b * c->sm_count;
  }
  if (const char *env = getenv("CRT_L2_PERSIST")) c->l2_persist = atoi(env);
  if (const char *env = getenv("CRT_TAIL_ITERS")) c->tail_iters = std::max(-1, atoi(env));
  if (const char *env = getenv("CRT_TAIL_CAP")) c->tail_cap = std::max(1, atoi(env));
  if (const char *env = getenv("CRT_TAIL_START")) c->tail_start = c->tail_start_closest = std::max(0, atoi(env));
  if (const char *env = getenv("CRT_TAIL_START_CLOSEST")) c->tail_start_closest = std::max(0, atoi(env));
  if (const char *env = getenv("CRT_TAIL_SMALL")) c->tail_small = std::max(0, atoi(env));
  if (const char *env = getenv("CRT_COOP_EARLY")) c->coop_early = std::max(0, std::min(CRT_COOP_MIN_BLOCKS, atoi(env)));
  c->l2_persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
  c->l2_window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
  if (c->l2_persist && c->l2_persist_max)
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_max);  // best effort: a refusal only loses the hint
  if (getenv("CRT_VERBOSE"))
    fprintf(stderr, "[crtb200] %s: %d SMs, L2 %d MiB, persisting L2 max %zu MiB, access-policy window max %zu MiB, persist %s\n", prop.name,
            prop.multiProcessorCount, prop.l2CacheSize >> 20, c->l2_persist_max >> 20, c->l2_window_max >> 20, c->l2_persist ? "on" : "off");
  *out = c;
  return CRTB200_OK;
}

const char *crtb200_last_error_ctx(const crtb200_ctx *c) { return c ? c->error.c_str() : g_error.c_str(); }

int crtb200_create_multi(const int *device_ids, int n, crtb200_ctx **out) {
  if (!out) return fail(CRTB200_ERR_ARG, "out is null");
  *out = nullptr;
  if (!device_ids || n < 1 || n > 64) return fail(CRTB200_ERR_ARG, "device list is null or its length is not in 1..64");
  crtb200_ctx *c = nullptr;
  int rc = crtb200_create(device_ids[0], &c);
  if (rc) return rc;
  for (int i = 1; i < n; i++) {
    crtb200_ctx *p = nullptr;
    rc = crtb200_create(device_ids[i], &p);
    if (rc == CRTB200_OK && p->device != c->device) {
      // the peer's kernels store into the primary's frame: map the primary's memory into the peer's address space
      int can = 0;
      cudaSetDevice(p->device);
      cudaError_t e = cudaDeviceCanAccessPeer(&can, p->device, c->device);
      if (e == cudaSuccess && !can) e = cudaErrorPeerAccessUnsupported;
      if (e == cudaSuccess) e = cudaDeviceEnablePeerAccess(c->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        e = cudaSuccess;
      }
      if (e != cudaSuccess)
        rc = fail(CRTB200_ERR_CUDA, std::string("peer access from device ") + std::to_string(p->device) + " to device " +
                                        std::to_string(c->device) + ": " + cudaGetErrorString(e));
    }
    if (rc == CRTB200_OK && cudaEventCreateWithFlags(&p->multi_done, cudaEventDisableTiming) != cudaSuccess)
      rc = fail(CRTB200_ERR_CUDA, "cudaEventCreate failed");
    if (rc) {
      if (p) crtb200_destroy(p);
      crtb200_destroy(c);
      return rc;
    }
    c->peers.push_back(p);
    c->workers.emplace_back(new Worker());
  }
  c->worker_rc.assign(c->peers.size(), 0);
  c->worker_msg.assign(c->peers.size(), std::string());
  cudaSetDevice(c->device);
  if (!c->peers.empty() && cudaEventCreateWithFlags(&c->multi_fork, cudaEventDisableTiming) != cudaSuccess) {
    crtb200_destroy(c);
    return fail(CRTB200_ERR_CUDA, "cudaEventCreate failed");
  }
  *out = c;
  return CRTB200_OK;
}

int crtb200_device_list(const crtb200_ctx *c, int *device_ids, int capacity, int *count) {
  if (!c || !count) return fail(CRTB200_ERR_ARG, "null argument");
  *count = 1 + (int)c->peers.size();
  if (device_ids)
    for (int i = 0; i < capacity && i < *count; i++) device_ids[i] = i == 0 ? c->device : c->peers[i - 1]->device;
  return CRTB200_OK;
}

int crtb200_destroy(crtb200_ctx *c) {
  if (!c) return CRTB200_OK;
  c->workers.clear();  // joins the helper threads
  for (crtb200_ctx *p : c->peers) crtb200_destroy(p);
  c->peers.clear();
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->multi_fork) cudaEventDestroy(c->multi_fork);
  if (c->multi_done) cudaEventDestroy(c->multi_done);
  c->arena.release(); c->vtx_normal.release(); c->top_refs.release();
  c->tri_shade.release(); c->vtx_uv.release(); c->meshes.release(); c->materials.release(); c->textures.release();
  c->texels.release(); c->lights.release(); c->frame.release(); c->frame8.release(); c->hits.release();
  c->mask.release(); c->frame_b.release(); c->frame8_b.release();
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (int k = 0; k < 2; k++) {
    if (c->rendered[k]) cudaEventDestroy(c->rendered[k]);
    if (c->copied[k]) cudaEventDestroy(c->copied[k]);
  }
  for (auto &q : c->sets) {
    q.release();
    if (q.stream) cudaStreamDestroy(q.stream);
    if (q.done) cudaEventDestroy(q.done);
  }
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  for (auto &q : c->sets) {
    if (q.coop_stream) cudaStreamDestroy(q.coop_stream);
    if (q.coop_fork) cudaEventDestroy(q.coop_fork);
    if (q.coop_join) cudaEventDestroy(q.coop_join);
  }
  if (c->band_stream) cudaStreamDestroy(c->band_stream);
  if (c->band_done) cudaEventDestroy(c->band_done);
  for (cudaEvent_t e : c->chunk_done) cudaEventDestroy(e);
  for (cudaEvent_t e : c->dbg_ev) cudaEventDestroy(e);
  c->stats_dev.release();
  for (auto &e : c->ev)
    if (e) cudaEventDestroy(e);
  for (auto &e : c->kev) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return CRTB200_OK;
}

int crtb200_set_concurrency(crtb200_ctx *c, uint32_t chunks_in_flight) {
  ErrScope scope(c);
  if (!c) return fail(CRTB200_ERR_ARG, "ctx is null");
  if (chunks_in_flight < 1 || chunks_in_flight > 16) return fail(CRTB200_ERR_ARG, "concurrency must be 1..16");
  c->concurrency = chunks_in_flight;
  c->cap_items = 0;
  for (crtb200_ctx *p : c->peers) crtb200_set_concurrency(p, chunks_in_flight);
  return CRTB200_OK;
}

int crtb200_set_queue_budget(crtb200_ctx *c, uint64_t bytes) {
  ErrScope scope(c);
  if (!c) return fail(CRTB200_ERR_ARG, "ctx is null");
  if (bytes < (64ull << 20)) return fail(CRTB200_ERR_ARG, "queue budget must be at least 64 MiB");
  c->queue_budget = bytes;
  c->cap_items = 0;
  for (crtb200_ctx *p : c->peers) crtb200_set_queue_budget(p, bytes);
  return CRTB200_OK;
}

// ---- host flattener H1: reference-numbered tree -> visiting-order, skip-linked nodes -------------------------
// The reference pops child[1] before child[0] (it pushes 0 then 1 on a std::stack, KDTree.cpp:65-72), visits every
// node whose box passes, and never reorders.  So the visiting order is a fixed total order of the nodes and a failed
// slab test simply jumps over the node's subtree.
static bool relayout_tree(const crtb200_kdnode *nodes, uint32_t n, uint32_t out_base, uint32_t ref_base,
                          uint32_t ref_count, std::vector<float4> &out, std::string &err, uint32_t *max_depth = nullptr) {
  if (max_depth) *max_depth = 0;
  if (n == 0) return true;
  std::vector<uint32_t> order;
  order.reserve(n);
  std::vector<uint32_t> stack;
  std::vector<uint8_t> seen(n, 0);
  stack.push_back(0);
  while (!stack.empty()) {
    uint32_t i = stack.back();
    stack.pop_back();
    if (i >= n) {
      err = "kd node child index out of range";
      return false;
    }
    if (seen[i]) {
      err = "kd tree is not a tree (node reachable twice)";
      return false;
    }
    seen[i] = 1;
    order.push_back(i);
    const crtb200_kdnode &nd = nodes[i];
    if (nd.leaf_count == 0) {
      if (nd.child[0] != CRTB200_INVALID) stack.push_back(nd.child[0]);
      if (nd.child[1] != CRTB200_INVALID) stack.push_back(nd.child[1]);
    } else if ((uint64_t)nd.leaf_start + nd.leaf_count > ref_count) {
      err = "kd leaf reference range out of bounds";
      return false;
    }
  }
  const uint32_t m = (uint32_t)order.size();
  if (max_depth) {  // order is a pre-order: a parent precedes its children
    std::vector<uint32_t> depth(n, 1);
    for (uint32_t k = 0; k < m; k++) {
      const crtb200_kdnode &nd = nodes[order[k]];
      *max_depth = std::max(*max_depth, depth[order[k]]);
      if (nd.leaf_count == 0)
        for (int side = 0; side < 2; side++)
          if (nd.child[side] != CRTB200_INVALID) depth[nd.child[side]] = depth[order[k]] + 1;
    }
  }
  std::vector<uint32_t> newidx(n, 0), size(n, 1);
  for (uint32_t k = 0; k < m; k++) newidx[order[k]] = k;
  for (uint32_t k = m; k-- > 0;) {
    const crtb200_kdnode &nd = nodes[order[k]];
    uint32_t s = 1;
    if (nd.leaf_count == 0) {
      if (nd.child[0] != CRTB200_INVALID) s += size[nd.child[0]];
      if (nd.child[1] != CRTB200_INVALID) s += size[nd.child[1]];
    }
    size[order[k]] = s;
  }
  out.resize(2 * (size_t)(out_base + m));
  for (uint32_t k = 0; k < m; k++) {
    const crtb200_kdnode &nd = nodes[order[k]];
    uint32_t a, b;
    if (nd.leaf_count) {
      a = CRT_LEAF_FLAG | nd.leaf_count;
      b = ref_base + nd.leaf_start;
    } else {
      a = out_base + k + size[order[k]];
      // b = the child visited second (child[0] when both exist): k_coop pushes both children of a passing node at once
      b = (nd.child[0] != CRTB200_INVALID && nd.child[1] != CRTB200_INVALID) ? out_base + newidx[nd.child[0]] : CRT_INVALID;
    }
    float4 lo = make_float4(nd.box_min[0], nd.box_min[1], nd.box_min[2], 0.f);
    float4 hi = make_float4(nd.box_max[0], nd.box_max[1], nd.box_max[2], 0.f);
    std::memcpy(&lo.w, &a, 4);
    std::memcpy(&hi.w, &b, 4);
    out[2 * (size_t)(out_base + k)] = lo;
    out[2 * (size_t)(out_base + k) + 1] = hi;
  }
  return true;
}

static bool box_nested(const crtb200_kdnode &parent, const crtb200_kdnode &child) {
  for (int k = 0; k < 3; k++)
    if (!(child.box_min[k] >= parent.box_min[k]) || !(child.box_max[k] <= parent.box_max[k]) || !(child.box_min[k] <= child.box_max[k]))
      return false;
  return true;
}

// ---- culling margin of one mesh (crt_device.cuh "Conservative culling") -----------------------
// mu = overhang + 4 * slop + 64 ulp * |coordinates|, in double:
//   overhang  how far the bounding box of a triangle sticks out of a leaf that lists it (max over all leaf references):
//             then bbox(T) lies inside L inflated by overhang for EVERY leaf L through which the reference can meet T
//   slop      how far outside T a point can lie and still pass Triangle::pointIsInTriangle's -FLT_EPSILON tests
//             (Triangle.cpp:37-57): eps * (1 + 8 e^2) * perimeter / (2 * area), e = longest edge (rounding of the edge
//             functions included); x4 for safety
//   the ulp term covers the rounding of the hit point and of the slab parameters for this mesh's extent (the ray
//   origin's share is added per ray, cull_margin_for)
// +inf (never cull) when the tree does not nest or a triangle's uploaded normal is not its geometric unit normal (then
// the slop bound does not describe what the triangle test accepts).  Zero-area triangles with a zero / non-finite
// normal are harmless: they never yield a finite t (Ray.cpp:11-18).
static float mesh_cull_margin(const crtb200_scene *s, const crtb200_mesh &me, bool nested) {
  const float inf = std::numeric_limits<float>::infinity();
  if (!nested) return inf;
  double overhang = 0.0, slop = 0.0, absmax = 0.0;
  const crtb200_kdnode *nodes = s->mesh_nodes + me.first_node;
  for (uint32_t k = 0; k < me.n_nodes; k++) {
    const crtb200_kdnode &nd = nodes[k];
    for (int i = 0; i < 3; i++) absmax = std::max(absmax, std::max(std::fabs((double)nd.box_min[i]), std::fabs((double)nd.box_max[i])));
    for (uint32_t q = 0; q < nd.leaf_count; q++) {
      const uint32_t *iv = s->triangle_vertex + 3 * (size_t)(me.first_triangle + s->mesh_leaf_refs[me.first_leaf_ref + nd.leaf_start + q]);
      for (int i = 0; i < 3; i++) {
        const double x0 = s->vertex_position[3 * (size_t)iv[0] + i], x1 = s->vertex_position[3 * (size_t)iv[1] + i],
                     x2 = s->vertex_position[3 * (size_t)iv[2] + i];
        const double lo = std::min(x0, std::min(x1, x2)), hi = std::max(x0, std::max(x1, x2));
        overhang = std::max(overhang, std::max((double)nd.box_min[i] - lo, hi - (double)nd.box_max[i]));
      }
    }
  }
  for (uint32_t t = 0; t < me.n_triangles; t++) {
    const uint32_t *iv = s->triangle_vertex + 3 * (size_t)(me.first_triangle + t);
    double p[3][3];
    for (int v = 0; v < 3; v++)
      for (int i = 0; i < 3; i++) {
        p[v][i] = s->vertex_position[3 * (size_t)iv[v] + i];
        absmax = std::max(absmax, std::fabs(p[v][i]));
      }
    double len[3];
    for (int k = 0; k < 3; k++) {
      const double *a = p[k], *b = p[(k + 1) % 3];
      len[k] = std::sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]));
    }
    const double u[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
    const double w[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
    const double cr[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    const double a2 = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    const float *nn = s->triangle_normal + 3 * (size_t)(me.first_triangle + t);
    if (!(a2 > 0.0) || !std::isfinite(a2)) {
      for (int i = 0; i < 3; i++)
        if (std::isfinite(nn[i]) && nn[i] != 0.0f) return inf;
      continue;
    }
    for (int i = 0; i < 3; i++)
      if (!(std::fabs((double)nn[i] - cr[i] / a2) <= 1e-3)) return inf;
    const double emax = std::max(len[0], std::max(len[1], len[2]));
    slop = std::max(slop, (double)CRT_FLT_EPSILON * (1.0 + 8.0 * emax * emax) * (len[0] + len[1] + len[2]) / a2);
  }
  const double mu = overhang + 4.0 * slop + 64.0 * (double)CRT_FLT_EPSILON * absmax;
  return (std::isfinite(mu) && mu < 1e30) ? (float)(mu * (1.0 + 1e-6)) : inf;
}

}  // extern "C"

static int upload_one(crtb200_ctx *c, const crtb200_scene *s);

// runs fn(peer index, peer) on every peer's helper thread and fn0 on the calling thread; first error wins
template <typename F0, typename F>
static int on_all_devices(crtb200_ctx *c, F0 fn0, F fn) {
  for (size_t i = 0; i < c->peers.size(); i++) {
    c->workers[i]->post([c, i, &fn] {
      ErrScope scope(c->peers[i]);
      c->worker_rc[i] = fn(i, c->peers[i]);
      c->worker_msg[i] = c->worker_rc[i] ? g_error : std::string();
    });
  }
  int rc = fn0();
  for (size_t i = 0; i < c->peers.size(); i++) {
    c->workers[i]->wait();
    if (rc == CRTB200_OK && c->worker_rc[i]) rc = fail(c->worker_rc[i], "device " + std::to_string(c->peers[i]->device) + ": " + c->worker_msg[i]);
  }
  return rc;
}

extern "C" {

int crtb200_upload_scene(crtb200_ctx *c, const crtb200_scene *s) {
  if (!c || !s) return fail(CRTB200_ERR_ARG, "null argument");
  ErrScope scope(c);
  if (c->peers.empty()) return upload_one(c, s);
  // scene replicated on every GPU; the host relayout runs once per GPU, in parallel
  return on_all_devices(c, [&] { return upload_one(c, s); }, [&](size_t, crtb200_ctx *p) { return upload_one(p, s); });
}

}  // extern "C"

static int upload_one(crtb200_ctx *c, const crtb200_scene *s) {
  if (s->abi_version != CRTB200_ABI_VERSION) return fail(CRTB200_ERR_ARG, "crtb200_scene.abi_version mismatch");
  if (s->width == 0 || s->height == 0) return fail(CRTB200_ERR_SCENE, "image size is zero");
  if ((uint64_t)s->width * s->height > 0x7FFFFFFFull / 4) return fail(CRTB200_ERR_SCENE, "image too large");
  if ((s->n_vertices && (!s->vertex_position || !s->vertex_normal)) || (s->n_triangles && (!s->triangle_vertex || !s->triangle_normal)) ||
      (s->n_meshes && !s->meshes) || (s->n_materials && !s->materials) || (s->n_lights && !s->lights) ||
      (s->n_textures && !s->textures) || (s->n_mesh_nodes && !s->mesh_nodes) || (s->n_top_nodes && !s->top_nodes) ||
      (s->n_mesh_leaf_refs && !s->mesh_leaf_refs) || (s->n_top_leaf_refs && !s->top_leaf_refs))
    return fail(CRTB200_ERR_SCENE, "a non-empty scene array is null");
  if (s->n_triangles >= 0x7FFFFFFFu || s->n_mesh_nodes >= 0x7FFFFFF0u) return fail(CRTB200_ERR_SCENE, "scene too large for 31-bit indices");
  CUDA_TRY(cudaSetDevice(c->device));
  c->have_scene = false;

  // triangles
  std::vector<float4> geom(3 * (size_t)s->n_triangles);
  std::vector<uint4> shade(s->n_triangles);
  std::vector<uint32_t> tri_mesh(s->n_triangles, CRTB200_INVALID);
  for (uint32_t m = 0; m < s->n_meshes; m++) {
    const crtb200_mesh &me = s->meshes[m];
    if ((uint64_t)me.first_triangle + me.n_triangles > s->n_triangles) return fail(CRTB200_ERR_SCENE, "mesh triangle range out of bounds");
    if (me.material >= s->n_materials) return fail(CRTB200_ERR_SCENE, "mesh material index out of range");
    if ((uint64_t)me.first_node + me.n_nodes > s->n_mesh_nodes) return fail(CRTB200_ERR_SCENE, "mesh node range out of bounds");
    if ((uint64_t)me.first_leaf_ref + me.n_leaf_refs > s->n_mesh_leaf_refs) return fail(CRTB200_ERR_SCENE, "mesh leaf-ref range out of bounds");
    for (uint32_t t = 0; t < me.n_triangles; t++) tri_mesh[me.first_triangle + t] = m;
  }
  for (uint32_t t = 0; t < s->n_triangles; t++) {
    const uint32_t *iv = s->triangle_vertex + 3 * (size_t)t;
    if (iv[0] >= s->n_vertices || iv[1] >= s->n_vertices || iv[2] >= s->n_vertices) return fail(CRTB200_ERR_SCENE, "triangle vertex index out of range");
    if (tri_mesh[t] == CRTB200_INVALID) return fail(CRTB200_ERR_SCENE, "triangle not owned by any mesh");
    const float *n = s->triangle_normal + 3 * (size_t)t;
    for (int k = 0; k < 3; k++) {
      const float *p = s->vertex_position + 3 * (size_t)iv[k];
      geom[3 * (size_t)t + k] = make_float4(p[0], p[1], p[2], n[k]);
    }
    shade[t] = make_uint4(iv[0], iv[1], iv[2], tri_mesh[t]);
  }
  // trees: mesh trees first, top-level tree last, one node array
  std::vector<float4> nodes;
  bool nested_ok = true;
  std::vector<uint32_t> refs(s->n_mesh_leaf_refs);
  std::vector<DMesh> meshes(s->n_meshes);
  uint32_t node_cursor = 0;
  std::string err;
  for (uint32_t m = 0; m < s->n_meshes; m++) {
    const crtb200_mesh &me = s->meshes[m];
    for (uint32_t k = 0; k < me.n_leaf_refs; k++) {
      uint32_t local = s->mesh_leaf_refs[me.first_leaf_ref + k];
      if (local >= me.n_triangles) return fail(CRTB200_ERR_SCENE, "mesh leaf reference out of range");
      refs[me.first_leaf_ref + k] = me.first_triangle + local;
    }
    const size_t before = nodes.size() / 2;
    uint32_t depth = 0;
    if (!relayout_tree(s->mesh_nodes + me.first_node, me.n_nodes, node_cursor, me.first_leaf_ref, me.n_leaf_refs, nodes, err, &depth))
      return fail(CRTB200_ERR_SCENE, err);
    if (depth > 40) nested_ok = false;  // k_coop's LIFOs are sized for trees up to this deep (the reference's limit is 26 levels)
    const uint32_t placed = (uint32_t)(nodes.size() / 2 - before);
    meshes[m].node_begin = node_cursor;
    meshes[m].node_end = node_cursor + placed;
    meshes[m].material = me.material;
    meshes[m].first_triangle = me.first_triangle;
    meshes[m].cull_margin = std::numeric_limits<float>::infinity();
    node_cursor += placed;
    for (uint32_t k = 0; k < me.n_nodes && nested_ok; k++) {  // relayout_tree has validated the child indices
      const crtb200_kdnode &p = s->mesh_nodes[me.first_node + k];
      for (int k3 = 0; k3 < 3; k3++)
        if (!(p.box_min[k3] <= p.box_max[k3])) nested_ok = false;
      if (p.leaf_count) continue;
      for (int side = 0; side < 2; side++)
        if (p.child[side] != CRTB200_INVALID && !box_nested(p, s->mesh_nodes[me.first_node + p.child[side]])) nested_ok = false;
    }
  }
  for (uint32_t k = 0; k < s->n_top_leaf_refs; k++)
    if (s->top_leaf_refs[k] >= s->n_meshes) return fail(CRTB200_ERR_SCENE, "top-level leaf reference out of range");
  const uint32_t top_begin = node_cursor;
  {
    const size_t before = nodes.size() / 2;
    if (!relayout_tree(s->top_nodes, s->n_top_nodes, node_cursor, 0, s->n_top_leaf_refs, nodes, err)) return fail(CRTB200_ERR_SCENE, err);
    node_cursor += (uint32_t)(nodes.size() / 2 - before);
  }
  const uint32_t top_end = node_cursor;
  // culling margins (one pass over the leaf references and the triangles of each mesh; meshes in parallel would be easy,
  // but this is 0.1 s per million triangles)
  for (uint32_t m = 0; m < s->n_meshes; m++) meshes[m].cull_margin = mesh_cull_margin(s, s->meshes[m], nested_ok);

  std::vector<float4> vn(s->n_vertices);
  for (uint32_t v = 0; v < s->n_vertices; v++)
    vn[v] = make_float4(s->vertex_normal[3 * (size_t)v], s->vertex_normal[3 * (size_t)v + 1], s->vertex_normal[3 * (size_t)v + 2], 0.f);
  std::vector<float2> uv;
  if (s->vertex_uv) {
    uv.resize(s->n_vertices);
    for (uint32_t v = 0; v < s->n_vertices; v++) uv[v] = make_float2(s->vertex_uv[3 * (size_t)v], s->vertex_uv[3 * (size_t)v + 1]);
  }
  std::vector<DMaterial> mats(s->n_materials);
  c->has_reflective = c->has_refractive = false;
  for (uint32_t m = 0; m < s->n_materials; m++) {
    const crtb200_material &mm = s->materials[m];
    if (mm.texture != CRTB200_INVALID && mm.texture >= s->n_textures) return fail(CRTB200_ERR_SCENE, "material texture index out of range");
    mats[m].type = mm.type;
    mats[m].smooth = mm.smooth_shading;
    mats[m].texture = mm.texture;
    mats[m].ior = mm.ior;
    for (int k = 0; k < 3; k++) mats[m].albedo[k] = mm.albedo[k];
    mats[m].pad = 0.f;
  }
  for (uint32_t m = 0; m < s->n_meshes; m++) {
    const uint32_t type = s->materials[s->meshes[m].material].type;
    if (type == CRTB200_MAT_REFLECTIVE) c->has_reflective = true;
    if (type == CRTB200_MAT_REFRACTIVE) c->has_refractive = true;
  }
  std::vector<DTexture> texs(s->n_textures);
  for (uint32_t t = 0; t < s->n_textures; t++) {
    const crtb200_texture &tt = s->textures[t];
    texs[t].kind = tt.kind;
    for (int k = 0; k < 3; k++) {
      texs[t].color_a[k] = tt.color_a[k];
      texs[t].color_b[k] = tt.color_b[k];
    }
    texs[t].scalar = tt.scalar;
    texs[t].width = tt.width;
    texs[t].height = tt.height;
    texs[t].texel_offset = tt.texel_offset;
    if (tt.kind == CRTB200_TEX_BITMAP) {
      if (tt.width == 0 || tt.height == 0 || tt.texel_offset + (uint64_t)tt.width * tt.height > s->n_texels || !s->texels)
        return fail(CRTB200_ERR_SCENE, "bitmap texture texel range out of bounds");
    }
  }
  std::vector<float> texels(s->texels ? s->texels : nullptr, s->texels ? s->texels + 3 * s->n_texels : nullptr);
  std::vector<DLight> lights(s->n_lights);
  for (uint32_t l = 0; l < s->n_lights; l++) {
    for (int k = 0; k < 3; k++) lights[l].pos[k] = s->lights[l].position[k];
    lights[l].intensity = static_cast<float>(s->lights[l].intensity);  // RayTracer.cpp:320
  }
  std::vector<uint32_t> top_refs(s->top_leaf_refs, s->top_leaf_refs + s->n_top_leaf_refs);

  {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b_nodes = al(nodes.size() * sizeof(float4)), b_refs = al(refs.size() * sizeof(uint32_t)),
                 b_geom = al(geom.size() * sizeof(float4));
    c->arena_used = b_nodes + b_refs + b_geom;
    c->nodes_bytes = b_nodes;
    CUDA_TRY(c->arena.ensure(std::max<size_t>(c->arena_used, 256)));
    c->nodes_p = reinterpret_cast<float4 *>(c->arena.p);
    c->leaf_refs_p = reinterpret_cast<uint32_t *>(c->arena.p + b_nodes);
    c->tri_geom_p = reinterpret_cast<float4 *>(c->arena.p + b_nodes + b_refs);
    if (!nodes.empty()) CUDA_TRY(cudaMemcpy(c->nodes_p, nodes.data(), nodes.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if (!refs.empty()) CUDA_TRY(cudaMemcpy(c->leaf_refs_p, refs.data(), refs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (!geom.empty()) CUDA_TRY(cudaMemcpy(c->tri_geom_p, geom.data(), geom.size() * sizeof(float4), cudaMemcpyHostToDevice));
  }
  c->nested_ok = nested_ok;
  CUDA_TRY(c->top_refs.upload(top_refs));
  CUDA_TRY(c->tri_shade.upload(shade));
  CUDA_TRY(c->vtx_normal.upload(vn));
  if (!uv.empty()) CUDA_TRY(c->vtx_uv.upload(uv));
  CUDA_TRY(c->meshes.upload(meshes));
  CUDA_TRY(c->materials.upload(mats));
  CUDA_TRY(c->textures.upload(texs));
  CUDA_TRY(c->texels.upload(texels));
  CUDA_TRY(c->lights.upload(lights));
  c->scene_bytes = nodes.size() * 16 + refs.size() * 4 + geom.size() * 16 + shade.size() * 16 + vn.size() * 16;

  DScene &d = c->sc;
  d.nodes = c->nodes_p;
  d.leaf_refs = c->leaf_refs_p;
  d.top_refs = c->top_refs.p;
  d.tri_geom = c->tri_geom_p;
  d.tri_shade = c->tri_shade.p;
  d.vtx_normal = c->vtx_normal.p;
  d.vtx_uv = uv.empty() ? nullptr : c->vtx_uv.p;
  d.meshes = c->meshes.p;
  d.materials = c->materials.p;
  d.textures = c->textures.p;
  d.texels = c->texels.p;
  d.lights = c->lights.p;
  d.n_lights = s->n_lights;
  d.n_meshes = s->n_meshes;
  d.top_begin = top_begin;
  d.top_end = top_end;
  d.width = s->width;
  d.height = s->height;
  for (int k = 0; k < 3; k++) d.bg[k] = s->background[k];
  d.dedup_meshes = s->n_meshes <= 64 ? 1u : (s->n_meshes <= 512 ? 2u : 0u);
  d.dedup_words = d.dedup_meshes == 2u ? (s->n_meshes + 31u) / 32u : 0u;

  const size_t px = (size_t)s->width * s->height;
  CUDA_TRY(c->frame.ensure(px * 3));
  CUDA_TRY(cudaMemset(c->frame.p, 0, px * 3 * sizeof(float)));  // colorBuffer starts at (0,0,0), RayTracer.cpp:47-50
  CUDA_TRY(c->frame8.ensure(px * 3));
  CUDA_TRY(cudaMemset(c->frame8.p, 0, px * 3));
  CUDA_TRY(c->stats_dev.ensure(48));
  c->mask_valid = false;
  c->cap_items = 0;
  c->have_scene = true;
  return CRTB200_OK;
}

// ---- frame planning ------------------------------------------------------------------------------------------
static uint32_t branching_sum(const crtb200_ctx *c, uint32_t max_depth, uint64_t *per_level) {
  // worst-case rays per level-0 item at each level (refractive hits spawn two children, RayTracer.cpp:398-412)
  uint64_t sum = 0;
  for (uint32_t l = 0; l <= max_depth; l++) {
    uint64_t b = 1;
    if (l > 0) {
      if (c->has_refractive)
        b = 1ull << std::min<uint32_t>(l, 40);
      else if (c->has_reflective)
        b = 1;
      else
        b = 0;
    }
    per_level[l] = b;
    sum += b;
  }
  return (uint32_t)std::min<uint64_t>(sum, 0xFFFFFFFFull);
}

// Marks the scene arena as L2-persisting for the kernels of `st` (one access-policy window per stream).  When the arena
// is larger than the persisting carve-out, hitRatio makes a matching fraction of its lines persist instead of thrashing.
static void apply_l2_window(crtb200_ctx *c, cudaStream_t st) {
  if (!c->l2_persist || !c->l2_persist_max || !c->l2_window_max || !c->arena_used) return;
  cudaStreamAttrValue v{};
  const size_t bytes = std::min(c->l2_persist == 3 ? c->nodes_bytes : c->arena_used, c->l2_window_max);
  v.accessPolicyWindow.base_ptr = c->arena.p;
  v.accessPolicyWindow.num_bytes = bytes;
  v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)c->l2_persist_max / (double)bytes);
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = c->l2_persist == 1 ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
  cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
  cudaGetLastError();  // a hint: never an error of the render
}

static int plan_queues(crtb200_ctx *c, uint32_t shard_items, uint32_t max_depth, uint32_t row_items, bool pipelined) {
  uint64_t per_level[CRT_MAX_LEVELS] = {0};
  branching_sum(c, max_depth, per_level);
  uint64_t sum = 0;
  for (uint32_t l = 0; l <= max_depth; l++) sum += per_level[l];
  const uint32_t n_lights = std::max<uint32_t>(1, c->sc.n_lights);
  const uint64_t bytes_per_node = 32 + 8 + 16 + 16 + 48 + n_lights;  // ray + hit + colour + comb + diffuse item + visibility bytes
  // sets used: up to `concurrency`, but never chunks smaller than 64 Ki items (launch overhead would dominate)
  uint32_t n_sets = std::max<uint32_t>(1, std::min<uint32_t>(c->concurrency, (shard_items + 65535u) / 65536u));
  // host-bound frames: chunks of at least ~512 Ki items (smaller ones only add launches and tails)
  if (pipelined) n_sets = std::max<uint32_t>(1, std::min<uint32_t>(n_sets, (shard_items + (1u << 19) - 1u) >> 19));
  // measured (profiles/r1_tuning.md): overlapping chunks only pays when band copies to the host ride along; the
  // persistent kernels already fill the GPU, extra chunks just add launches and tails
  if (!pipelined) {
    n_sets = 1;
    if (const char *env = getenv("CRT_DEVICE_CHUNKS")) n_sets = std::max(1, std::min<int>(atoi(env), (int)std::min<uint32_t>(16u, (shard_items + 65535u) / 65536u)));
  }
  uint64_t items = c->queue_budget / (bytes_per_node * sum * n_sets);
  items &= ~31ull;
  if (items < 32 * 64) return fail(CRTB200_ERR_MEMORY, "queue budget too small for one chunk at this ray depth");
  // host-bound frames: one chunk per stream at 1080p, two at 4K (a band's copy overlaps the traversal of the next band;
  // every chunk pays its own kernel tails and, in reflective scenes, the latency of every level again)
  uint32_t per_set = (pipelined && n_sets > 1 && shard_items >= (1u << 22)) ? 2u : 1u;
  if (const char *env = getenv("CRT_HOST_CHUNKS_PER_SET")) per_set = (uint32_t)std::max(1, std::min(8, atoi(env)));  // tools: e2e tuning
  const uint32_t parts = n_sets * per_set;
  uint64_t even = ((uint64_t)shard_items + parts - 1) / parts;
  even = ((even + row_items - 1) / row_items) * row_items;
  items = std::min<uint64_t>(items, even);
  // node ids and (diffuse item, light) slots are 32-bit on the device
  if (items * sum * n_lights >= 0x7FFFFFFFull) items = ((0x7FFFFFFFull / (sum * n_lights)) - 32) & ~31ull;
  if (items < 32) return fail(CRTB200_ERR_MEMORY, "too many lights x ray-tree nodes for 32-bit queue slots");
  if (items >= row_items) items = (items / row_items) * row_items;  // whole tile rows (band copies need it)
  c->n_chunks = (uint32_t)((shard_items + items - 1) / items);
  if (c->cap_items == items && c->cap_depth == max_depth && c->cap_sets == n_sets) return CRTB200_OK;
  if (c->sets.size() < n_sets) c->sets.resize(n_sets);
  if (!c->fork_ev) CUDA_TRY(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
  for (uint32_t k = 0; k < n_sets; k++) {
    crtb200_ctx::QueueSet &q = c->sets[k];
    if (!q.stream) CUDA_TRY(cudaStreamCreateWithFlags(&q.stream, cudaStreamNonBlocking));
    apply_l2_window(c, q.stream);
    if (!q.done) CUDA_TRY(cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming));
    uint64_t total = 0;
    for (uint32_t l = 0; l <= max_depth; l++) {
      q.lv.offset[l] = (uint32_t)total;
      total += per_level[l] * items;
    }
    for (uint32_t l = max_depth + 1; l <= CRT_MAX_LEVELS; l++) q.lv.offset[l] = (uint32_t)total;
    const uint64_t secondary = total - items;
    CUDA_TRY(q.ray_o.ensure(std::max<uint64_t>(secondary, 1)));
    CUDA_TRY(q.ray_d.ensure(std::max<uint64_t>(secondary, 1)));
    CUDA_TRY(q.hit_tri.ensure(total));
    CUDA_TRY(q.hit_t.ensure(total));
    CUDA_TRY(q.color.ensure(total));
    CUDA_TRY(q.comb.ensure(total));
    CUDA_TRY(q.dq.ensure(3 * total));
    CUDA_TRY(q.vis.ensure(std::max<uint64_t>(1, total * std::max<uint32_t>(1, c->sc.n_lights))));
    // hand-off records: every lane of a traversal grid hands off at most once per launch
    const uint64_t ovf_cap = (uint64_t)std::max(c->blocks_closest, c->blocks_shadow) * CRT_TRAV_BLOCK;
    CUDA_TRY(q.ovf.ensure(3 * ovf_cap));
    // every record starts out "not published" (r0.x = CRT_INVALID); k_coop leaves a taken record in that state again
    CUDA_TRY(cudaMemsetAsync(q.ovf.p, 0xFF, 3 * ovf_cap * sizeof(uint4), q.stream));
    if (!q.coop_stream) {
      int lo = 0, hi = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least priority: the traversal kernel's blocks go first
      CUDA_TRY(cudaStreamCreateWithPriority(&q.coop_stream, cudaStreamNonBlocking, lo));
      CUDA_TRY(cudaEventCreateWithFlags(&q.coop_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&q.coop_join, cudaEventDisableTiming));
    }
    // counts | work cursors | hand-off counters; every cursor / counter group on a 128-byte line of its own
    const size_t n_counts = 64, n_work = CRT_CTL_STRIDE * (CRT_MAX_LEVELS + 2), n_ovf = CRT_CTL_STRIDE * (CRT_MAX_LEVELS + 1);
    CUDA_TRY(q.ctl.ensure(n_counts + n_work + n_ovf));
    q.work = q.ctl.p + n_counts;
    q.lv.ovf = q.ovf.p;
    q.lv.ovf_ctl = q.ctl.p + n_counts + n_work;
    q.lv.ovf_cap = (uint32_t)std::min<uint64_t>(ovf_cap, (uint64_t)c->tail_cap * c->blocks_coop * CRT_COOP_WARPS);
    q.lv.tail_iters = 0;
    q.lv.tail_start = 512;
    q.lv.tail_start_closest = 256;
    q.lv.tail_small = 0;
    q.lv.skip_zero_terms = 0;
    q.lv.ray_o = q.ray_o.p;
    q.lv.ray_d = q.ray_d.p;
    q.lv.hit_tri = q.hit_tri.p;
    q.lv.hit_t = q.hit_t.p;
    q.lv.color = q.color.p;
    q.lv.comb = q.comb.p;
    q.lv.dq = q.dq.p;
    q.lv.vis = q.vis.p;
    q.lv.counts = q.ctl.p;
    q.lv.stats = c->stats_dev.p;
  }
  c->cap_items = (uint32_t)items;
  c->cap_depth = max_depth;
  c->cap_sets = n_sets;
  return CRTB200_OK;
}

// Pixel coverage of the rectangle list; the mask is only needed when the rectangles do not tile the image exactly
// (the reference then leaves pixels unrendered).
static int plan_mask(crtb200_ctx *c, const crtb200_options *o) {
  const uint32_t W = c->sc.width, H = c->sc.height;
  std::vector<crtb200_rect> rects(o->rects, o->rects + o->n_rects);
  if (c->mask_valid && rects.size() == c->mask_rects.size() &&
      (rects.empty() || !std::memcmp(rects.data(), c->mask_rects.data(), rects.size() * sizeof(crtb200_rect))))
    return CRTB200_OK;
  c->mask_rects = rects;
  c->mask_needed = false;
  if (!rects.empty()) {
    std::vector<uint8_t> m((size_t)W * H, 0);
    for (const auto &r : rects) {
      const uint32_t r1 = std::min<uint64_t>(H, (uint64_t)r.row + r.height), c1 = std::min<uint64_t>(W, (uint64_t)r.col + r.width);
      for (uint32_t y = r.row; y < r1; y++) std::memset(&m[(size_t)y * W + r.col], 1, r.col < c1 ? c1 - r.col : 0);
    }
    size_t covered = 0;
    for (uint8_t v : m) covered += v;
    if (covered != m.size()) {
      c->mask_needed = true;
      CUDA_TRY(c->mask.ensure(m.size()));
      CUDA_TRY(cudaMemcpy(c->mask.p, m.data(), m.size(), cudaMemcpyHostToDevice));
    }
  }
  c->mask_valid = true;
  return CRTB200_OK;
}

static cudaEvent_t next_event(crtb200_ctx *c) {
  if (c->kev_used == c->kev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->kev.push_back(e);
  }
  return c->kev[c->kev_used++];
}

// dynamic shared memory of the kernels that walk one ray per lane: the visited-mesh bitset of scenes with > 64 meshes
static size_t dyn_smem(const crtb200_ctx *c, int block) { return (size_t)c->sc.dedup_words * (size_t)block * sizeof(uint32_t); }

template <bool COUNT, bool CULL, bool WIDE>
static void launch_closest_w(crtb200_ctx *c, bool primary, const Frame &fr, const Levels &lv, uint32_t level, uint32_t *work,
                             cudaStream_t st) {
  if (primary)
    k_closest<true, COUNT, CRT_REFILL_CLOSEST, CRT_LOOP_MODE, CULL, WIDE><<<c->blocks_closest, CRT_TRAV_BLOCK, dyn_smem(c, CRT_TRAV_BLOCK), st>>>(c->sc, fr, lv, level, work);
  else
    k_closest<false, COUNT, CRT_REFILL_CLOSEST, CRT_LOOP_MODE, CULL, WIDE><<<c->blocks_closest_sec, CRT_TRAV_BLOCK, dyn_smem(c, CRT_TRAV_BLOCK), st>>>(c->sc, fr, lv, level, work);
}
// the WIDE flavour (visited-mesh set in shared memory) only for scenes that need it, and never when counting the
// reference's visit-all work (no de-duplication there)
template <bool COUNT, bool CULL>
static void launch_closest(crtb200_ctx *c, bool primary, const Frame &fr, const Levels &lv, uint32_t level, uint32_t *work,
                           cudaStream_t st) {
  if (!COUNT && c->sc.dedup_meshes == 2u)
    launch_closest_w<COUNT, CULL, !COUNT>(c, primary, fr, lv, level, work, st);
  else
    launch_closest_w<COUNT, CULL, false>(c, primary, fr, lv, level, work, st);
}
template <int COUNT, bool CULL>
static void launch_shadow(crtb200_ctx *c, const Frame &fr, const Levels &lv, uint32_t *work, cudaStream_t st) {
  if (COUNT != 1 && c->sc.dedup_meshes == 2u)
    k_shadow<COUNT, CRT_REFILL_SHADOW, CRT_LOOP_MODE, CULL, (COUNT != 1)><<<c->blocks_shadow, CRT_TRAV_BLOCK, dyn_smem(c, CRT_TRAV_BLOCK), st>>>(c->sc, fr, lv, work);
  else
    k_shadow<COUNT, CRT_REFILL_SHADOW, CRT_LOOP_MODE, CULL, false><<<c->blocks_shadow, CRT_TRAV_BLOCK, 0, st>>>(c->sc, fr, lv, work);
}

// k_coop around one traversal launch.  coop_begin forks the side stream BEFORE the traversal kernel is
// launched; coop_end marks the traversal kernel's end behind it, starts the early pass on the side stream (its blocks
// get SM resources as the traversal kernel's blocks exit), joins, and runs the final pass over what is left.
static bool coop_early_on(const crtb200_ctx *c, bool secondary_level) {
  return c->coop_early > 0 && c->cap_sets == 1 && !secondary_level;
}
static int coop_begin(crtb200_ctx *c, crtb200_ctx::QueueSet &q, bool secondary_level) {
  if (!coop_early_on(c, secondary_level)) return CRTB200_OK;
  CUDA_TRY(cudaEventRecord(q.coop_fork, q.stream));
  CUDA_TRY(cudaStreamWaitEvent(q.coop_stream, q.coop_fork, 0));
  return CRTB200_OK;
}
template <bool SHADOW, bool CULL>
static int coop_end(crtb200_ctx *c, crtb200_ctx::QueueSet &q, bool primary, const Frame &fr, uint32_t level, uint32_t &launches) {
  const uint32_t launch = SHADOW ? (uint32_t)CRT_MAX_LEVELS : level;
  auto run = [&](int blocks, cudaStream_t st, uint32_t early) {
    if (SHADOW)
      k_coop<true, false, CULL, CRT_COOP_GROUP><<<blocks, 32 * CRT_COOP_WARPS, 0, st>>>(c->sc, fr, q.lv, 0, early);
    else if (primary)
      k_coop<false, true, CULL, CRT_COOP_GROUP><<<blocks, 32 * CRT_COOP_WARPS, 0, st>>>(c->sc, fr, q.lv, level, early);
    else
      k_coop<false, false, CULL, CRT_COOP_GROUP><<<blocks, 32 * CRT_COOP_WARPS, 0, st>>>(c->sc, fr, q.lv, level, early);
    launches++;
  };
  if (coop_early_on(c, !SHADOW && !primary)) {
    k_mark<<<1, 32, 0, q.stream>>>(q.lv.ovf_ctl + CRT_CTL_STRIDE * launch + 24u);
    run(c->sm_count * c->coop_early, q.coop_stream, 1u);
    CUDA_TRY(cudaEventRecord(q.coop_join, q.coop_stream));
    CUDA_TRY(cudaStreamWaitEvent(q.stream, q.coop_join, 0));
    launches++;
  }
  run(c->blocks_coop, q.stream, 0u);
  return CRTB200_OK;
}

// Host destinations of crtb200_render: each chunk's band of rows is copied back on the chunk's own stream right after
// its k_store, so the device->host copy of band k overlaps the traversal of the other chunks.
struct HostOut {
  float *rgb = nullptr;
  uint8_t *rgb8 = nullptr;
  crtb200_hit *hits = nullptr;
};

// Enqueue one frame on `st`.  d_rgb / d_rgb8 / d_hits / d_slab are device pointers (any may be null).
static int enqueue_frame(crtb200_ctx *c, const crtb200_camera *cam, const crtb200_options *o, float *d_rgb,
                         uint8_t *d_rgb8, HitRec *d_hits, float *d_slab, cudaStream_t st, bool timed,
                         const HostOut *host = nullptr, bool *host_done = nullptr, bool batch_continuation = false) {
  if (!c->have_scene) return fail(CRTB200_ERR_STATE, "no scene uploaded");
  if (o->max_depth > 31) return fail(CRTB200_ERR_ARG, "max_depth > 31 is not supported");
  if (o->n_rects && !o->rects) return fail(CRTB200_ERR_ARG, "n_rects > 0 but rects is null");
  if (o->traversal > 1) return fail(CRTB200_ERR_ARG, "unknown traversal mode");
  if (o->count_work > 2) return fail(CRTB200_ERR_ARG, "unknown count_work mode");
  // traversal 0 (default): conservative culling + tail hand-off, both exact and both resting on the nesting property of
  // the uploaded trees; traversal 1 and the visit-all counting mode walk the reference's literal itinerary
  const bool cull = o->traversal == 0 && o->count_work != 1 && c->nested_ok;
  const bool handoff = o->traversal == 0 && o->count_work == 0 && c->nested_ok && c->tail_iters >= 0;
  const uint32_t shard_count = o->shard_count ? o->shard_count : 1;
  if (o->shard_index >= shard_count) return fail(CRTB200_ERR_ARG, "shard_index >= shard_count");
  int rc = plan_mask(c, o);
  if (rc) return rc;
  const uint32_t W = c->sc.width, H = c->sc.height;
  Frame fr{};
  for (int k = 0; k < 3; k++) fr.cam.pos[k] = cam->position[k];
  for (int k = 0; k < 9; k++) fr.cam.rot[k] = cam->rotation[k];
  fr.tiles_x = (W + 7) / 8;
  fr.n_tiles = fr.tiles_x * ((H + 3) / 4);
  fr.shard_index = o->shard_index;
  fr.shard_count = shard_count;
  fr.mask = c->mask_needed ? c->mask.p : nullptr;
  fr.max_depth = o->max_depth;
  fr.resolve = (c->has_reflective || c->has_refractive) ? 1u : 0u;
  fr.shadow_bias = o->shadow_bias;
  fr.reflection_bias = o->reflection_bias;
  fr.refraction_bias = o->refraction_bias;
  const uint32_t shard_tiles = (fr.n_tiles > o->shard_index) ? (fr.n_tiles - o->shard_index + shard_count - 1) / shard_count : 0;
  const uint32_t shard_items = shard_tiles * 32u;
  // unsharded frames: chunks are whole rows of tiles, so a chunk owns a contiguous band of image rows
  const uint32_t row_items = (shard_count == 1) ? fr.tiles_x * 32u : 32u;
  rc = plan_queues(c, std::max(shard_items, 32u), o->max_depth, row_items, host != nullptr);
  if (rc) return rc;
  const bool secondary = c->has_reflective || c->has_refractive;
  const uint32_t levels = secondary ? o->max_depth + 1 : 1;
  // Band copies need chunks that own whole rows of tiles: with a small queue budget or a deep refractive scene a chunk
  // is smaller than a tile row, two chunks on different streams would then share a 4-row band and a band copy could
  // pick up the other chunk's stale pixels.  Then the caller copies the whole frame after the join instead.
  const bool band_copies = host && shard_count == 1 && c->cap_items % row_items == 0;
  if (host_done) *host_done = band_copies;
  const int grid_simple = c->sm_count * 8;

  // frames after the first of a batch (crtb200_render_frames) keep accumulating into the same counters and time span
  if (!batch_continuation) CUDA_TRY(cudaMemsetAsync(c->stats_dev.p, 0, 48 * sizeof(unsigned long long), st));
#if CRT_PHASE_CLOCKS
  CUDA_TRY(cudaMemsetAsync(c->stats_dev.p + 27, 0xFF, sizeof(unsigned long long), st));  // atomicMin slots
  CUDA_TRY(cudaMemsetAsync(c->stats_dev.p + 31, 0xFF, sizeof(unsigned long long), st));
#endif
  c->kev_used = 0;
  c->kev_kind.clear();
  if (timed && !batch_continuation) CUDA_TRY(cudaEventRecord(c->ev[0], st));
  // fork: every set's stream waits for the caller's stream, chunks go round-robin, the caller's stream joins at the end.
  // Per-kernel event pairs are only recorded without concurrency (overlapping kernels would inflate each other).
  const uint32_t n_sets = c->cap_sets;
  const bool per_kernel = timed && n_sets == 1 && !batch_continuation;
  CUDA_TRY(cudaEventRecord(c->fork_ev, st));
  for (uint32_t k = 0; k < n_sets; k++) CUDA_TRY(cudaStreamWaitEvent(c->sets[k].stream, c->fork_ev, 0));
  uint32_t launches = 0, chunk = 0;
  // band copies leave on their own stream, so a set starts its next chunk while its last band is still in flight
  const bool band_stream = band_copies && n_sets > 1 && !getenv("CRT_NO_BAND_STREAM");
  if (band_stream) {
    if (!c->band_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->band_stream, cudaStreamNonBlocking));
    if (!c->band_done) CUDA_TRY(cudaEventCreateWithFlags(&c->band_done, cudaEventDisableTiming));
    while (c->chunk_done.size() < c->n_chunks) {
      cudaEvent_t e = nullptr;
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->chunk_done.push_back(e);
    }
  }
  for (uint32_t begin = 0; begin < shard_items; begin += c->cap_items, chunk++) {
    crtb200_ctx::QueueSet &q = c->sets[chunk % n_sets];
    cudaStream_t qs = q.stream;
    fr.item_begin = begin;
    fr.n_items0 = std::min(c->cap_items, shard_items - begin);
    CUDA_TRY(cudaMemsetAsync(q.ctl.p, 0, q.ctl.n * sizeof(uint32_t), qs));
    q.lv.tail_iters = handoff ? (uint32_t)(c->tail_iters + 1) : 0u;
    q.lv.skip_zero_terms = (o->traversal == 0 && o->count_work != 1) ? 1u : 0u;
    q.lv.tail_start = (uint32_t)c->tail_start;
    q.lv.tail_start_closest = (uint32_t)c->tail_start_closest;
    q.lv.tail_small = (uint32_t)c->tail_small;
    for (uint32_t l = 0; l < levels; l++) {
      if (per_kernel) {
        cudaEventRecord(next_event(c), qs);
        c->kev_kind.push_back(0);
      }
      if (handoff) {
        rc = coop_begin(c, q, l != 0);
        if (rc) return rc;
      }
      if (o->count_work && cull)
        launch_closest<true, true>(c, l == 0, fr, q.lv, l, q.work + CRT_CTL_STRIDE * l, qs);
      else if (o->count_work)
        launch_closest<true, false>(c, l == 0, fr, q.lv, l, q.work + CRT_CTL_STRIDE * l, qs);
      else if (cull)
        launch_closest<false, true>(c, l == 0, fr, q.lv, l, q.work + CRT_CTL_STRIDE * l, qs);
      else
        launch_closest<false, false>(c, l == 0, fr, q.lv, l, q.work + CRT_CTL_STRIDE * l, qs);
      if (handoff) {
        if (per_kernel) {
          cudaEventRecord(next_event(c), qs);
          cudaEventRecord(next_event(c), qs);
          c->kev_kind.push_back(2);
        }
        rc = cull ? coop_end<false, true>(c, q, l == 0, fr, l, launches) : coop_end<false, false>(c, q, l == 0, fr, l, launches);
        if (rc) return rc;
      }
      if (per_kernel) cudaEventRecord(next_event(c), qs);
      k_shade<<<grid_simple, 256, 0, qs>>>(c->sc, fr, q.lv, l);
      launches += 2;
    }
    if (per_kernel) {
      cudaEventRecord(next_event(c), qs);
      c->kev_kind.push_back(1);
    }
    uint32_t *swork = q.work + CRT_CTL_STRIDE * CRT_MAX_LEVELS;
    if (handoff) {
      rc = coop_begin(c, q, false);
      if (rc) return rc;
    }
    if (o->count_work == 1)
      launch_shadow<1, false>(c, fr, q.lv, swork, qs);
    else if (o->count_work == 2 && cull)
      launch_shadow<2, true>(c, fr, q.lv, swork, qs);
    else if (o->count_work == 2)
      launch_shadow<2, false>(c, fr, q.lv, swork, qs);
    else if (cull)
      launch_shadow<0, true>(c, fr, q.lv, swork, qs);
    else
      launch_shadow<0, false>(c, fr, q.lv, swork, qs);
    if (handoff) {
      if (per_kernel) {
        cudaEventRecord(next_event(c), qs);
        cudaEventRecord(next_event(c), qs);
        c->kev_kind.push_back(3);
      }
      rc = cull ? coop_end<true, true>(c, q, false, fr, 0, launches) : coop_end<true, false>(c, q, false, fr, 0, launches);
      if (rc) return rc;
    }
    if (per_kernel) cudaEventRecord(next_event(c), qs);
    k_accumulate<<<grid_simple, 256, 0, qs>>>(c->sc, fr, q.lv);
    launches += 2;
    for (uint32_t l = levels - 1; l-- > 0;) {
      k_resolve<<<grid_simple, 256, 0, qs>>>(c->sc, fr, q.lv, l);
      launches++;
    }
    k_store<<<grid_simple, 256, 0, qs>>>(c->sc, fr, q.lv, d_rgb, d_rgb8, d_hits, d_slab);
    launches++;
    if (band_copies) {
      const uint32_t row0 = (begin / row_items) * 4u;
      const uint32_t row1 = std::min<uint32_t>(H, ((begin + fr.n_items0 + row_items - 1) / row_items) * 4u);
      const size_t off = (size_t)row0 * W, cnt = (size_t)(row1 - row0) * W;
      cudaStream_t cs = qs;
      if (band_stream) {
        CUDA_TRY(cudaEventRecord(c->chunk_done[chunk], qs));
        CUDA_TRY(cudaStreamWaitEvent(c->band_stream, c->chunk_done[chunk], 0));
        cs = c->band_stream;
      }
      if (host->rgb) CUDA_TRY(cudaMemcpyAsync(host->rgb + off * 3, d_rgb + off * 3, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, cs));
      if (host->rgb8) CUDA_TRY(cudaMemcpyAsync(host->rgb8 + off * 3, d_rgb8 + off * 3, cnt * 3, cudaMemcpyDeviceToHost, cs));
      if (host->hits) CUDA_TRY(cudaMemcpyAsync(host->hits + off, d_hits + off, cnt * sizeof(HitRec), cudaMemcpyDeviceToHost, cs));
      if (getenv("CRT_CHUNK_TIMES")) {
        while (c->dbg_ev.size() < 2 * (size_t)(chunk + 1)) {
          cudaEvent_t e = nullptr;
          CUDA_TRY(cudaEventCreate(&e));
          c->dbg_ev.push_back(e);
        }
        CUDA_TRY(cudaEventRecord(c->dbg_ev[2 * chunk], qs));
        CUDA_TRY(cudaEventRecord(c->dbg_ev[2 * chunk + 1], cs));
        c->dbg_chunks = chunk + 1;
      }
    }
  }
  if (band_stream) {
    CUDA_TRY(cudaEventRecord(c->band_done, c->band_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, c->band_done, 0));
  }
  for (uint32_t k = 0; k < n_sets; k++) {
    CUDA_TRY(cudaEventRecord(c->sets[k].done, c->sets[k].stream));
    CUDA_TRY(cudaStreamWaitEvent(st, c->sets[k].done, 0));
  }
  if (timed) CUDA_TRY(cudaEventRecord(c->ev[1], st));
  CUDA_TRY(cudaGetLastError());
  c->last.kernel_launches = launches;
  c->last.levels = levels;
  return CRTB200_OK;
}

// Multi-GPU frame (crtb200_create_multi): shard i of n goes to GPU i; every GPU stores into the SAME destination
// buffers (device memory of the primary, peer-mapped on the others).  The peers start after `st` has reached this point
// (earlier consumers of the frame on the primary's stream are done) and `st` continues after every peer has stored.
static int enqueue_any(crtb200_ctx *c, const crtb200_camera *cam, const crtb200_options *o, float *d_rgb, uint8_t *d_rgb8,
                       HitRec *d_hits, float *d_slab, cudaStream_t st, bool timed, const HostOut *host = nullptr,
                       bool *host_done = nullptr, bool batch_continuation = false) {
  if (c->peers.empty() || o->shard_count > 1)
    return enqueue_frame(c, cam, o, d_rgb, d_rgb8, d_hits, d_slab, st, timed, host, host_done, batch_continuation);
  if (host_done) *host_done = false;  // the caller copies the assembled frame after the join
  const uint32_t n = 1u + (uint32_t)c->peers.size();
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaEventRecord(c->multi_fork, st));
  crtb200_options o0 = *o;
  o0.shard_index = 0;
  o0.shard_count = n;
  int rc = on_all_devices(
      c, [&] { return enqueue_frame(c, cam, &o0, d_rgb, d_rgb8, d_hits, nullptr, st, timed, nullptr, nullptr, batch_continuation); },
      [&](size_t i, crtb200_ctx *p) {
        crtb200_options oi = *o;
        oi.shard_index = (uint32_t)i + 1u;
        oi.shard_count = n;
        CUDA_TRY(cudaSetDevice(p->device));
        CUDA_TRY(cudaStreamWaitEvent(p->stream, c->multi_fork, 0));
        int r = enqueue_frame(p, cam, &oi, d_rgb, d_rgb8, d_hits, nullptr, p->stream, true, nullptr, nullptr, batch_continuation);
        if (r) return r;
        p->last_pending = true;
        CUDA_TRY(cudaEventRecord(p->multi_done, p->stream));
        return (int)CRTB200_OK;
      });
  CUDA_TRY(cudaSetDevice(c->device));
  if (rc) return rc;
  for (crtb200_ctx *p : c->peers) CUDA_TRY(cudaStreamWaitEvent(st, p->multi_done, 0));
  if (timed) CUDA_TRY(cudaEventRecord(c->ev[1], st));  // the frame ends when the last GPU has stored its shard
  return CRTB200_OK;
}

static int collect_stats_one(crtb200_ctx *c, bool timed);
// counters of the primary plus those of the peers (rays, tests, launches); times are the primary's (fork to join)
static int collect_stats(crtb200_ctx *c, bool timed) {
  int rc = collect_stats_one(c, timed);
  if (rc) return rc;
  for (crtb200_ctx *p : c->peers) {
    if (!p->last_pending) continue;
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    rc = collect_stats_one(p, false);
    p->last_pending = false;
    if (rc) return rc;
    c->last.rays_primary += p->last.rays_primary;
    c->last.rays_shadow += p->last.rays_shadow;
    c->last.rays_reflection += p->last.rays_reflection;
    c->last.rays_refraction += p->last.rays_refraction;
    c->last.node_tests_closest += p->last.node_tests_closest;
    c->last.triangle_tests_closest += p->last.triangle_tests_closest;
    c->last.node_tests_shadow += p->last.node_tests_shadow;
    c->last.triangle_tests_shadow += p->last.triangle_tests_shadow;
    c->last.handoff_closest += p->last.handoff_closest;
    c->last.handoff_shadow += p->last.handoff_shadow;
    c->last.shadow_rays_zero_term += p->last.shadow_rays_zero_term;
    c->last.kernel_launches += p->last.kernel_launches;
  }
  CUDA_TRY(cudaSetDevice(c->device));
  return CRTB200_OK;
}

static int collect_stats_one(crtb200_ctx *c, bool timed) {
  unsigned long long st[48];
  CUDA_TRY(cudaMemcpy(st, c->stats_dev.p, sizeof(st), cudaMemcpyDeviceToHost));
#if CRT_PHASE_CLOCKS
  for (int k = 0; k < 2; k++) {
    const unsigned long long *p = st + 8 + 8 * k;
    const double tot = (double)(p[0] + p[1] + p[2] + p[3] + p[4]);
    fprintf(stderr, "[phase clocks] %s: refill %.1f%% slow %.1f%% node %.1f%% tri %.1f%% other %.1f%% | node iterations %llu, tri phases (MODE 2) / ranges stolen (k_*_s) %llu, rounds %llu\n",
            k ? "k_shadow " : "k_closest", 100.0 * p[0] / tot, 100.0 * p[1] / tot, 100.0 * p[2] / tot, 100.0 * p[3] / tot, 100.0 * p[4] / tot,
            p[5], p[6], p[7]);
    {
      unsigned long long hist[2][32];
      cudaMemcpyFromSymbol(hist, g_iter_hist, sizeof(hist));
      fprintf(stderr, "[phase clocks] %s rays by log2(node iterations + 1) (cumulative over frames):", k ? "k_shadow " : "k_closest");
      for (int b = 0; b < 20; b++) fprintf(stderr, " %llu", hist[k][b]);
      fprintf(stderr, "\n");
    }
    {
      unsigned long long tail[2][2];
      cudaMemcpyFromSymbol(tail, g_tail, sizeof(tail));
      const double clk = (double)(p[0] + p[1] + p[2] + p[3] + p[4]);
      fprintf(stderr, "[phase clocks] %s after the queue ran dry (cumulative): %.1f%% of warp time, %.1f of 32 lanes busy on average\n", k ? "k_shadow " : "k_closest",
              100.0 * (double)tail[k][0] / clk, tail[k][0] ? (double)tail[k][1] / (double)tail[k][0] : 0.0);
    }
    if (getenv("CRT_WARP_DUMP")) {
      static unsigned long long rec[2][8192][6];
      cudaMemcpyFromSymbol(rec, g_warp_rec, sizeof(rec));
      unsigned long long t0 = ~0ull;
      for (int w = 0; w < 8192; w++) if (rec[k][w][1] && rec[k][w][0] < t0) t0 = rec[k][w][0];
      std::vector<std::pair<unsigned long long, int>> ends;
      for (int w = 0; w < 8192; w++) if (rec[k][w][1]) ends.push_back({rec[k][w][1] - t0, w});
      std::sort(ends.begin(), ends.end());
      fprintf(stderr, "[warp dump] %s: %zu warps; end-time percentiles (us): p10 %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f\n", k ? "k_shadow " : "k_closest", ends.size(),
              ends[ends.size() / 10].first * 1e-3, ends[ends.size() / 2].first * 1e-3, ends[ends.size() * 9 / 10].first * 1e-3,
              ends[ends.size() * 99 / 100].first * 1e-3, ends.back().first * 1e-3);
      for (size_t i = ends.size() >= 6 ? ends.size() - 6 : 0; i < ends.size(); i++) {
        const unsigned long long *r = rec[k][ends[i].second];
        fprintf(stderr, "[warp dump]   warp %d: start %.0f us end %.0f us rounds %llu node-iterations %llu clk-after-dry %llu stolen %llu\n", ends[i].second,
                (r[0] - t0) * 1e-3, (r[1] - t0) * 1e-3, r[2], r[3], r[4], r[5]);
      }
    }
    const unsigned long long *g = st + 24 + 4 * k;  // sum of warp lifetimes, warps, last end, first start (globaltimer ns)
    if (g[1])
      fprintf(stderr, "[phase clocks] %s: %llu warps, mean warp lifetime %.1f us, kernel span %.1f us (first warp start -> last warp end; one launch per frame only)\n",
              k ? "k_shadow " : "k_closest", g[1], 1e-3 * (double)g[0] / (double)g[1], 1e-3 * (double)(g[2] - g[3]));
  }
#endif
  c->last.rays_primary = st[0];
  c->last.rays_shadow = st[1];
  c->last.rays_reflection = st[2];
  c->last.rays_refraction = st[3];
  c->last.node_tests_closest = st[4];
  c->last.triangle_tests_closest = st[5];
  c->last.node_tests_shadow = st[6];
  c->last.triangle_tests_shadow = st[7];
  c->last.handoff_closest = st[32];
  c->last.handoff_shadow = st[33];
  c->last.shadow_rays_zero_term = st[40];
#if CRT_COOP_STATS
  fprintf(stderr, "[coop stats] closest: %llu walks, %llu iterations, %llu box tests, %llu triangle tests | shadow: %llu walks, %llu iterations, %llu box tests, %llu triangle tests\n",
          st[32], st[34], st[35], st[36], st[33], st[37], st[38], st[39]);
#endif
  if (timed) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    c->last.device_ms = ms;
    c->last.closest_ms = c->last.shadow_ms = c->last.coop_closest_ms = c->last.coop_shadow_ms = 0.0;
    for (size_t k = 0; k < c->kev_kind.size() && 2 * k + 1 < c->kev_used; k++) {
      float kms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&kms, c->kev[2 * k], c->kev[2 * k + 1]));
      // closest_ms / shadow_ms include their k_coop part, which is also reported on its own
      if (c->kev_kind[k] == 0 || c->kev_kind[k] == 2) c->last.closest_ms += kms;
      if (c->kev_kind[k] == 1 || c->kev_kind[k] == 3) c->last.shadow_ms += kms;
      if (c->kev_kind[k] == 2) c->last.coop_closest_ms += kms;
      if (c->kev_kind[k] == 3) c->last.coop_shadow_ms += kms;
      if (getenv("CRT_KERNEL_TIMES"))  // tools: one line per traversal launch of the frame, in launch order
        fprintf(stderr, "[kernel times] %s %.3f ms\n", c->kev_kind[k] == 0 ? "k_closest" : c->kev_kind[k] == 1 ? "k_shadow" : "k_coop", kms);
    }
  }
  return CRTB200_OK;
}

extern "C" {

int crtb200_render(crtb200_ctx *c, const crtb200_camera *cam, const crtb200_options *o, float *rgb_out,
                   uint8_t *rgb8_out, crtb200_hit *hits_out, crtb200_stats *stats) {
  ErrScope scope(c);
  if (!c || !cam || !o) return fail(CRTB200_ERR_ARG, "null argument");
  const auto t0 = std::chrono::steady_clock::now();
  CUDA_TRY(cudaSetDevice(c->device));
  if (!c->have_scene) return fail(CRTB200_ERR_STATE, "no scene uploaded");
  const size_t px = (size_t)c->sc.width * c->sc.height;
  if (hits_out) CUDA_TRY(c->hits.ensure(px));
  if (hits_out) CUDA_TRY(cudaMemsetAsync(c->hits.p, 0xFF, px * sizeof(HitRec), c->stream));
  c->last = crtb200_stats{};
  // Rectangle grids that leave pixels uncovered keep the previous frame there (colorBuffer persistence), so the whole
  // persistent frame is copied after the render; otherwise each chunk's band is copied as soon as it is stored.
  const bool shard = o->shard_count > 1;
  HostOut host;
  host.rgb = rgb_out;
  host.rgb8 = rgb8_out;
  host.hits = hits_out;
  int rc = plan_mask(c, o);
  if (rc) return rc;
  const bool want_bands = !shard && !c->mask_needed && (rgb_out || rgb8_out || hits_out);
  bool banded = false;
  rc = enqueue_any(c, cam, o, c->frame.p, rgb8_out ? c->frame8.p : nullptr, hits_out ? c->hits.p : nullptr, nullptr,
                   c->stream, true, want_bands ? &host : nullptr, &banded);
  if (rc) return rc;
  if (!banded) {
    if (rgb_out) CUDA_TRY(cudaMemcpyAsync(rgb_out, c->frame.p, px * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (rgb8_out) CUDA_TRY(cudaMemcpyAsync(rgb8_out, c->frame8.p, px * 3, cudaMemcpyDeviceToHost, c->stream));
    if (hits_out) CUDA_TRY(cudaMemcpyAsync(hits_out, c->hits.p, px * sizeof(HitRec), cudaMemcpyDeviceToHost, c->stream));
  }
  const double enqueue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const double sync_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  rc = collect_stats(c, true);
  if (rc) return rc;
  c->last.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (banded && c->dbg_chunks && getenv("CRT_CHUNK_TIMES")) {
    fprintf(stderr, "[chunk times] host: enqueue %.3f ms, synchronised %.3f ms, stats read %.3f ms; device frame %.3f ms\n", enqueue_ms, sync_ms,
            c->last.total_ms, c->last.device_ms);
    for (uint32_t k = 0; k < c->dbg_chunks; k++) {
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, c->ev[0], c->dbg_ev[2 * k]);
      cudaEventElapsedTime(&b, c->ev[0], c->dbg_ev[2 * k + 1]);
      fprintf(stderr, "[chunk times]   chunk %u (up to %u items, set %u): stored at %.3f ms, band on the host at %.3f ms\n", k, c->cap_items, k % c->cap_sets, a, b);
    }
    c->dbg_chunks = 0;
  }
  if (stats) *stats = c->last;
  return CRTB200_OK;
}

static void add_stats(crtb200_stats &total, const crtb200_stats &s) {
  total.rays_primary += s.rays_primary;
  total.rays_shadow += s.rays_shadow;
  total.rays_reflection += s.rays_reflection;
  total.rays_refraction += s.rays_refraction;
  total.node_tests_closest += s.node_tests_closest;
  total.triangle_tests_closest += s.triangle_tests_closest;
  total.node_tests_shadow += s.node_tests_shadow;
  total.triangle_tests_shadow += s.triangle_tests_shadow;
  total.handoff_closest += s.handoff_closest;
  total.handoff_shadow += s.handoff_shadow;
  total.shadow_rays_zero_term += s.shadow_rays_zero_term;
  total.device_ms += s.device_ms;
  total.closest_ms += s.closest_ms;
  total.shadow_ms += s.shadow_ms;
  total.coop_closest_ms += s.coop_closest_ms;
  total.coop_shadow_ms += s.coop_shadow_ms;
  total.kernel_launches += s.kernel_launches;
  total.levels = s.levels;
}

// The animation loop of app/animation.cpp:24-38 (setCamera + render per frame), batched and pipelined: no host
// synchronisation between frames, two frames in flight -- frame f + 1 is rendered while frame f's pixels travel to the
// host on a copy stream.  Rectangle lists that leave pixels uncovered need colorBuffer persistence from frame to frame
// (one buffer, in order), and tile shards write no full frame: those take the frame-by-frame path.
int crtb200_render_frames(crtb200_ctx *c, const crtb200_camera *cams, uint32_t n_frames, const crtb200_options *o,
                          float *rgb_out, uint8_t *rgb8_out, crtb200_stats *stats) {
  ErrScope scope(c);
  if (!c || !cams || !o) return fail(CRTB200_ERR_ARG, "null argument");
  if (!c->have_scene) return fail(CRTB200_ERR_STATE, "no scene uploaded");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t px = (size_t)c->sc.width * c->sc.height;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = plan_mask(c, o);
  if (rc) return rc;
  if (c->mask_needed || o->shard_count > 1 || o->count_work != 0 || n_frames < 2) {
    crtb200_stats total{};
    for (uint32_t f = 0; f < n_frames; f++) {
      crtb200_stats s{};
      rc = crtb200_render(c, cams + f, o, rgb_out ? rgb_out + f * px * 3 : nullptr, rgb8_out ? rgb8_out + f * px * 3 : nullptr, nullptr, &s);
      if (rc) return rc;
      add_stats(total, s);
    }
    total.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = total;
    return CRTB200_OK;
  }
  CUDA_TRY(c->frame_b.ensure(px * 3));
  if (rgb8_out) CUDA_TRY(c->frame8_b.ensure(px * 3));
  if (!c->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; k++) {
    if (!c->rendered[k]) CUDA_TRY(cudaEventCreateWithFlags(&c->rendered[k], cudaEventDisableTiming));
    if (!c->copied[k]) CUDA_TRY(cudaEventCreateWithFlags(&c->copied[k], cudaEventDisableTiming));
  }
  c->last = crtb200_stats{};
  uint32_t launches = 0;
  for (uint32_t f = 0; f < n_frames; f++) {
    const int k = (int)(f & 1u);
    float *d_rgb = k ? c->frame_b.p : c->frame.p;
    uint8_t *d_rgb8 = rgb8_out ? (k ? c->frame8_b.p : c->frame8.p) : nullptr;
    if (f >= 2) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->copied[k], 0));  // this pair's previous frame has left the device
    rc = enqueue_any(c, cams + f, o, d_rgb, d_rgb8, nullptr, nullptr, c->stream, true, nullptr, nullptr, f > 0);
    if (rc) return rc;
    launches += c->last.kernel_launches;
    CUDA_TRY(cudaEventRecord(c->rendered[k], c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->rendered[k], 0));
    if (rgb_out) CUDA_TRY(cudaMemcpyAsync(rgb_out + f * px * 3, d_rgb, px * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
    if (rgb8_out) CUDA_TRY(cudaMemcpyAsync(rgb8_out + f * px * 3, d_rgb8, px * 3, cudaMemcpyDeviceToHost, c->copy_stream));
    CUDA_TRY(cudaEventRecord(c->copied[k], c->copy_stream));
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
  if ((n_frames & 1u) == 0u) {  // the last frame went into the second pair: it is the persistent colour buffer now
    std::swap(c->frame, c->frame_b);
    std::swap(c->frame8, c->frame8_b);
  }
  rc = collect_stats(c, true);
  if (rc) return rc;
  c->last.kernel_launches = launches;
  c->last.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (stats) *stats = c->last;
  return CRTB200_OK;
}

int crtb200_render_device(crtb200_ctx *c, const crtb200_camera *cam, const crtb200_options *o, float *d_rgb_out,
                          uint8_t *d_rgb8_out, void *stream) {
  ErrScope scope(c);
  if (!c || !cam || !o) return fail(CRTB200_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  const uint32_t shard_count = o->shard_count ? o->shard_count : 1;
  c->last = crtb200_stats{};
  c->last_pending = true;
  if (shard_count > 1 && o->shard_full_frame)  // the shard's pixels at their place in a full frame (peer-mapped, crtb200_ipc_open)
    return enqueue_frame(c, cam, o, d_rgb_out, d_rgb8_out, nullptr, nullptr, (cudaStream_t)stream, true);
  if (shard_count > 1) {
    // sharded: d_rgb_out is the shard's compact slab (crtb200_shard_items x 3 floats), see crtb200_assemble_shards
    if (d_rgb8_out) return fail(CRTB200_ERR_ARG, "sharded rendering writes a float slab only");
    return enqueue_frame(c, cam, o, nullptr, nullptr, nullptr, d_rgb_out, (cudaStream_t)stream, true);
  }
  return enqueue_any(c, cam, o, d_rgb_out, d_rgb8_out, nullptr, nullptr, (cudaStream_t)stream, true);
}

int crtb200_last_stats(crtb200_ctx *c, crtb200_stats *stats) {
  ErrScope scope(c);
  if (!c || !stats) return fail(CRTB200_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  if (c->last_pending) {
    CUDA_TRY(cudaEventSynchronize(c->ev[1]));
    int rc = collect_stats(c, true);
    if (rc) return rc;
    c->last_pending = false;
  }
  *stats = c->last;
  return CRTB200_OK;
}

int crtb200_shard_items(crtb200_ctx *c, uint32_t shard_count, uint32_t *items) {
  ErrScope scope(c);
  if (!c || !items || shard_count == 0) return fail(CRTB200_ERR_ARG, "bad argument");
  if (!c->have_scene) return fail(CRTB200_ERR_STATE, "no scene uploaded");
  const uint32_t tiles = ((c->sc.width + 7) / 8) * ((c->sc.height + 3) / 4);
  *items = ((tiles + shard_count - 1) / shard_count) * 32u;
  return CRTB200_OK;
}

int crtb200_ipc_alloc(int device, size_t bytes, void **d_ptr_out) {
  if (!d_ptr_out || bytes == 0) return fail(CRTB200_ERR_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaMalloc(d_ptr_out, bytes));
  CUDA_TRY(cudaMemset(*d_ptr_out, 0, bytes));
  return CRTB200_OK;
}
int crtb200_ipc_free(int device, void *d_ptr) {
  if (!d_ptr) return CRTB200_OK;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaFree(d_ptr));
  return CRTB200_OK;
}
int crtb200_ipc_export(void *d_ptr, uint8_t handle_out[CRTB200_IPC_HANDLE_BYTES]) {
  if (!d_ptr || !handle_out) return fail(CRTB200_ERR_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == CRTB200_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, d_ptr));
  std::memcpy(handle_out, &h, sizeof(h));
  return CRTB200_OK;
}
int crtb200_ipc_open(int device, const uint8_t handle[CRTB200_IPC_HANDLE_BYTES], void **d_ptr_out) {
  if (!handle || !d_ptr_out) return fail(CRTB200_ERR_ARG, "null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return CRTB200_OK;
}
int crtb200_ipc_close(int device, void *d_ptr) {
  if (!d_ptr) return CRTB200_OK;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
  return CRTB200_OK;
}

int crtb200_assemble_shards(crtb200_ctx *c, const float *d_slabs, uint32_t shard_count, float *d_rgb_out,
                            uint8_t *d_rgb8_out, void *stream) {
  ErrScope scope(c);
  if (!c || !d_slabs || shard_count == 0) return fail(CRTB200_ERR_ARG, "bad argument");
  if (!c->have_scene) return fail(CRTB200_ERR_STATE, "no scene uploaded");
  CUDA_TRY(cudaSetDevice(c->device));
  uint32_t items = 0;
  crtb200_shard_items(c, shard_count, &items);
  Frame fr{};
  fr.tiles_x = (c->sc.width + 7) / 8;
  fr.n_tiles = fr.tiles_x * ((c->sc.height + 3) / 4);
  // rectangle lists that leave pixels uncovered: the assembled frame keeps what d_rgb_out held there (colorBuffer
  // persistence), exactly like the unsharded path -- the mask is the one of the last render's rectangle list
  fr.mask = c->mask_needed ? c->mask.p : nullptr;
  k_assemble<<<c->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(c->sc, fr, d_slabs, items, shard_count, d_rgb_out, d_rgb8_out);
  CUDA_TRY(cudaGetLastError());
  return CRTB200_OK;
}

int crtb200_generate_rays(crtb200_ctx *c, const crtb200_camera *cam, float *rays_out) {
  ErrScope scope(c);
  if (!c || !cam || !rays_out) return fail(CRTB200_ERR_ARG, "null argument");
  if (!c->have_scene) return fail(CRTB200_ERR_STATE, "no scene uploaded");
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t n = (size_t)c->sc.width * c->sc.height;
  DevBuf<float> d;
  CUDA_TRY(d.ensure(n * 6));
  DCamera dc;
  for (int k = 0; k < 3; k++) dc.pos[k] = cam->position[k];
  for (int k = 0; k < 9; k++) dc.rot[k] = cam->rotation[k];
  k_generate_rays<<<c->sm_count * 8, 256, 0, c->stream>>>(c->sc, dc, d.p);
  cudaError_t e = cudaMemcpyAsync(rays_out, d.p, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  d.release();
  CUDA_TRY(e);
  return CRTB200_OK;
}

int crtb200_debug_powf5(crtb200_ctx *c, const float *x, uint32_t n, float *out) {
  ErrScope scope(c);
  if (!c || !x || !out) return fail(CRTB200_ERR_ARG, "null argument");
  if (n == 0) return CRTB200_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  DevBuf<float> dx, dy;
  cudaError_t e = dx.ensure(n);
  if (e == cudaSuccess) e = dy.ensure(n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dx.p, x, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    k_powf5<<<c->sm_count * 8, 256, 0, c->stream>>>(dx.p, n, dy.p);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dy.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dx.release();
  dy.release();
  CUDA_TRY(e);
  return CRTB200_OK;
}

int crtb200_trace_rays(crtb200_ctx *c, const float *rays, uint32_t n, uint32_t ray_type, uint32_t traversal,
                       const float *max_distance, crtb200_hit *hits_out, uint8_t *occluded_out) {
  ErrScope scope(c);
  if (!c || !rays) return fail(CRTB200_ERR_ARG, "null argument");
  if (!c->have_scene) return fail(CRTB200_ERR_STATE, "no scene uploaded");
  if (ray_type > 3) return fail(CRTB200_ERR_ARG, "bad ray type");
  if (traversal > 1) return fail(CRTB200_ERR_ARG, "unknown traversal mode");  // caller-supplied rays always take the literal walk
  if (ray_type == CRTB200_RAY_SHADOW ? (!max_distance || !occluded_out) : !hits_out) return fail(CRTB200_ERR_ARG, "missing output / distance array");
  if (n == 0) return CRTB200_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  DevBuf<float> d_rays, d_dist;
  DevBuf<HitRec> d_hits;
  DevBuf<uint8_t> d_occ;
  cudaError_t e = d_rays.ensure((size_t)n * 6);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays.p, rays, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  if (ray_type == CRTB200_RAY_SHADOW) {
    if (e == cudaSuccess) e = d_dist.ensure(n);
    if (e == cudaSuccess) e = d_occ.ensure(n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dist.p, max_distance, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  } else if (e == cudaSuccess) {
    e = d_hits.ensure(n);
  }
  if (e == cudaSuccess) {
    k_query<<<c->sm_count * 8, 256, dyn_smem(c, 256), c->stream>>>(c->sc, d_rays.p, n, ray_type, d_dist.p, d_hits.p, d_occ.p);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    if (ray_type == CRTB200_RAY_SHADOW)
      e = cudaMemcpyAsync(occluded_out, d_occ.p, n, cudaMemcpyDeviceToHost, c->stream);
    else
      e = cudaMemcpyAsync(hits_out, d_hits.p, (size_t)n * sizeof(HitRec), cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  d_rays.release();
  d_dist.release();
  d_hits.release();
  d_occ.release();
  CUDA_TRY(e);
  return CRTB200_OK;
}

}  // extern "C"
