// odw_api.cu — host side of the C ABI declared in include/odw.h: device memory, scene/source upload,
// BVH build, kernel launches, result copy-out.  No CPU fallback: without a CUDA device every entry point
// that needs one fails with ODW_ENODEVICE / ODW_ECUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>
#include "odw_device.cuh"

extern "C" cudaError_t odw_launch_trace(const TraceParams* p, bool mc, bool bvh, int need, int blocks, size_t smem, cudaStream_t st);
extern "C" cudaError_t odw_launch_sample(const DSource* src, unsigned long long seed, unsigned long long first_ray,
                                         unsigned long long n, double* first_out, double* phi_out, double* origins,
                                         double* dirs, unsigned int* keys, float bound, int blocks, cudaStream_t st);
extern "C" cudaError_t odw_sort_pairs(void* temp, size_t* temp_bytes, const unsigned int* keys_in, unsigned int* keys_out,
                                      const unsigned int* vals_in, unsigned int* vals_out, unsigned int n, int begin_bit, int end_bit, cudaStream_t st);
extern "C" int odw_trace_occupancy(bool mc, bool bvh, int need, size_t smem);
extern "C" int odw_trace_threads(void);
// wavefront kernels (odw_wavefront.cu)
extern "C" size_t odw_wf_pool_bytes_per_ray(void);
extern "C" cudaError_t odw_wf_generate(const TraceParams* p, bool mc, void* pool, size_t cap, float bound, unsigned int n, cudaStream_t st);
extern "C" cudaError_t odw_wf_traverse(const TraceParams* p, void* pool, size_t cap, void* hits, unsigned int n,
                                       unsigned int* fetch_counter, const unsigned int* order, void* pool_ordered, int need, int blocks, cudaStream_t st);
extern "C" cudaError_t odw_wf_interact(const TraceParams* p, bool mc, void* pool_in, void* hits, void* pool_out, size_t cap, float bound,
                                       unsigned int n, unsigned int* n_next, int bounce, int keys, cudaStream_t st);
extern "C" cudaError_t odw_wf_iota(unsigned int* v, unsigned int n, cudaStream_t st);
extern "C" cudaError_t odw_wf_sort(void* temp, size_t* temp_bytes, void* pool, size_t cap, unsigned int* keys_out, const unsigned int* iota,
                                   unsigned int* order, unsigned int n, cudaStream_t st);
extern "C" cudaError_t odw_wf_tail(const TraceParams* p, bool mc, void* pool, size_t cap, unsigned int n, int bounce, cudaStream_t st);
extern "C" int odw_wf_traverse_occupancy(int n_nodes);
extern "C" cudaError_t odw_wf_traverse4(const TraceParams* p, void* pool, size_t cap, void* hits, unsigned int n, unsigned int* fetch_counter,
                                        const unsigned int* order, void* pool_ordered, int need, int regen, int blocks, cudaStream_t st);
extern "C" cudaError_t odw_wf_generate_keys(const TraceParams* p, void* pool, size_t cap, float bound, unsigned int n, cudaStream_t st);

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return fail(e_ == cudaErrorMemoryAllocation ? ODW_ENOMEM : ODW_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

static const int SMEM_FACE_LIMIT = 64;      // scenes up to this many faces are staged in shared memory and brute-forced

struct odw_engine {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // device->host copies of finished chunks (odw_trace_mc_host)
  static const int MAX_WAVE_STREAMS = 8, DEFAULT_WAVE_STREAMS = 4;
  cudaStream_t wave_stream[MAX_WAVE_STREAMS] = {};   // [0] = stream; launch waves rotate over them, so that the tail of one wave overlaps the head of the next
  static const uint64_t MAX_WAVES = 16384;
  unsigned long long* wave_counters = nullptr;   // [MAX_WAVES] device: next unclaimed ray of each launch wave of the request in flight
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_WAVE_STREAMS] = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_trace[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
  Counters* pinned_counters = nullptr;  // [3], page-locked: [0], [1] chunk counters of odw_trace_mc_host, [2] scratch of the wavefront loop
  std::string name;
  // size-keyed pool so that per-call result buffers are not re-allocated every step
  std::multimap<size_t, void*> pool;
  size_t pooled_bytes = 0;

  int alloc(void** out, size_t bytes) {
    if (bytes == 0) bytes = 16;
    auto it = pool.lower_bound(bytes);
    if (it != pool.end() && it->first <= bytes + bytes/4) {
      *out = it->second; pooled_bytes -= it->first; alloc_size[*out] = it->first; pool.erase(it);
      return ODW_OK;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {                 // release the pool and retry once
      cudaGetLastError();
      for (auto& kv : pool) cudaFree(kv.second);
      pool.clear(); pooled_bytes = 0;
      e = cudaMalloc(out, bytes);
      if (e != cudaSuccess) { cudaGetLastError(); return fail(ODW_ENOMEM, "cudaMalloc(" + std::to_string(bytes) + " bytes) failed"); }
    }
    alloc_size[*out] = bytes;
    return ODW_OK;
  }
  void release(void* p) {
    if (!p) return;
    auto it = alloc_size.find(p);
    if (it == alloc_size.end()) { cudaFree(p); return; }
    pool.emplace(it->second, p); pooled_bytes += it->second;
    alloc_size.erase(it);
  }
  std::map<void*, size_t> alloc_size;
};

struct odw_scene {
  odw_engine* eng = nullptr;
  DScene d{};
  std::vector<void*> owned;
  bool use_bvh = false;
  bool wavefront = true;                // BVH scenes: wavefront kernels (odw_wavefront.cu) instead of the register-resident kernel
  size_t smem = 0;
  int n_groups = 0;
  double extent = 0;                    // max |coordinate| over all face boxes
  bool ext_optics = false;              // a grating, a stochastic surface model, a finite absorption length or an even-asphere face somewhere (FEAT_EXT)
  std::vector<BvhNode2> bvh_host;       // un-widened device nodes (boxes rounded outward)
  std::vector<BvhNode2> bvh_staging;    // widened copy being uploaded
  BvhNode2* bvh_dev = nullptr;
  float bvh_margin = -1.0f;             // margin the device copy currently carries
  std::vector<Bvh4Node> bvh4_host, bvh4_staging;   // 4-wide tree of the wavefront traversal, same protocol
  Bvh4Node* bvh4_dev = nullptr;
};

struct odw_source {
  odw_engine* eng = nullptr;
  DSource d{};
  std::vector<void*> owned;
  double max_ray_length_scale = 1, max_intersections_scale = 1;
  double origin_bound = 0;              // max |coordinate| a ray origin of this source can have
  std::vector<int32_t> ignored;
};

struct odw_result {
  odw_engine* eng = nullptr;
  HitBuffers hb{};
  Counters* dcounters = nullptr;
  double* dbins = nullptr;
  DBinning* dbinnings = nullptr;
  std::vector<DBinning> binnings;
  size_t total_bins = 0;
  int32_t* d_nseg = nullptr; double* d_final_point = nullptr; double* d_final_power = nullptr; int32_t* d_final_medium = nullptr;
  double* d_in_o = nullptr; double* d_in_d = nullptr; double* d_in_p = nullptr;
  odw_counts counts{};
  uint64_t n_rays = 0;
  float ms = 0;
};

// ------------------------------------------------------------------------------------------
extern "C" int odw_abi_version(void) { return ODW_ABI_VERSION; }
extern "C" const char* odw_last_error(void) { return g_err.c_str(); }

extern "C" int odw_engine_create(int device_id, odw_engine** out) {
  if (!out) return fail(ODW_EINVAL, "odw_engine_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(ODW_ENODEVICE, std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                               "); this library has no CPU fallback");
  }
  if (device_id < 0 || device_id >= n) return fail(ODW_ENODEVICE, "device id " + std::to_string(device_id) + " out of range");
  CU(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device_id));
  odw_engine* eng = new odw_engine();
  eng->device = device_id;
  eng->sm_count = prop.multiProcessorCount;
  eng->name = prop.name;
  CU(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&eng->copy_stream, cudaStreamNonBlocking));
  eng->wave_stream[0] = eng->stream;
  for (int i = 1; i < odw_engine::MAX_WAVE_STREAMS; ++i) CU(cudaStreamCreateWithFlags(&eng->wave_stream[i], cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&eng->ev_fork, cudaEventDisableTiming));
  for (int i = 1; i < odw_engine::MAX_WAVE_STREAMS; ++i) CU(cudaEventCreateWithFlags(&eng->ev_join[i], cudaEventDisableTiming));
  CU(cudaEventCreate(&eng->ev0));
  CU(cudaEventCreate(&eng->ev1));
  for (int i = 0; i < 2; ++i) { CU(cudaEventCreateWithFlags(&eng->ev_trace[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&eng->ev_copy[i], cudaEventDisableTiming)); }
  CU(cudaHostAlloc((void**)&eng->pinned_counters, 3*sizeof(Counters), cudaHostAllocDefault));
  CU(cudaMalloc((void**)&eng->wave_counters, odw_engine::MAX_WAVES*sizeof(unsigned long long)));
  CU(cudaMemset(eng->wave_counters, 0, odw_engine::MAX_WAVES*sizeof(unsigned long long)));
  *out = eng;
  return ODW_OK;
}

extern "C" void odw_engine_destroy(odw_engine* eng) {
  if (!eng) return;
  cudaSetDevice(eng->device);
  for (auto& kv : eng->pool) cudaFree(kv.second);
  if (eng->ev0) cudaEventDestroy(eng->ev0);
  if (eng->ev1) cudaEventDestroy(eng->ev1);
  for (int i = 0; i < 2; ++i) { if (eng->ev_trace[i]) cudaEventDestroy(eng->ev_trace[i]); if (eng->ev_copy[i]) cudaEventDestroy(eng->ev_copy[i]); }
  if (eng->pinned_counters) cudaFreeHost(eng->pinned_counters);
  if (eng->wave_counters) cudaFree(eng->wave_counters);
  if (eng->copy_stream) cudaStreamDestroy(eng->copy_stream);
  for (int i = 1; i < odw_engine::MAX_WAVE_STREAMS; ++i) { if (eng->wave_stream[i]) cudaStreamDestroy(eng->wave_stream[i]); if (eng->ev_join[i]) cudaEventDestroy(eng->ev_join[i]); }
  if (eng->ev_fork) cudaEventDestroy(eng->ev_fork);
  if (eng->stream) cudaStreamDestroy(eng->stream);
  delete eng;
}

extern "C" int odw_engine_device_name(const odw_engine* eng, char* buf, int buflen) {
  if (!eng || !buf || buflen <= 0) return fail(ODW_EINVAL, "odw_engine_device_name: bad argument");
  snprintf(buf, (size_t)buflen, "%s (%d SMs)", eng->name.c_str(), eng->sm_count);
  return ODW_OK;
}

extern "C" int odw_engine_stream(const odw_engine* eng, void** stream_out) {
  if (!eng || !stream_out) return fail(ODW_EINVAL, "odw_engine_stream: bad argument");
  *stream_out = (void*)eng->stream;
  return ODW_OK;
}

extern "C" int odw_host_alloc(odw_engine* eng, uint64_t bytes, void** out) {
  if (!eng || !out) return fail(ODW_EINVAL, "odw_host_alloc: NULL argument");
  *out = nullptr;
  CU(cudaSetDevice(eng->device));
  cudaError_t e = cudaHostAlloc(out, std::max<uint64_t>(bytes, 16), cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return fail(ODW_ENOMEM, "cudaHostAlloc(" + std::to_string(bytes) + " bytes) failed"); }
  return ODW_OK;
}

extern "C" void odw_host_free(odw_engine* eng, void* p) {
  if (!eng || !p) return;
  cudaSetDevice(eng->device);
  cudaFreeHost(p);
}

template <typename T>
static int upload(odw_engine* eng, std::vector<void*>& owned, const T* host, size_t n, const T** dev) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(n, 1)*sizeof(T);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ODW_ENOMEM, "cudaMalloc failed in upload"); }
  owned.push_back(p);
  if (n) CU(cudaMemcpy(p, host, n*sizeof(T), cudaMemcpyHostToDevice));
  *dev = reinterpret_cast<const T*>(p);
  return ODW_OK;
}

// ---- SAH BVH over face boxes ---------------------------------------------------------------
namespace {
struct Box { double lo[3], hi[3];
  void reset() { for (int i = 0; i < 3; ++i) { lo[i] = 1e300; hi[i] = -1e300; } }
  void grow(const Box& b) { for (int i = 0; i < 3; ++i) { lo[i] = std::min(lo[i], b.lo[i]); hi[i] = std::max(hi[i], b.hi[i]); } }
  double area() const { double d[3] = { hi[0]-lo[0], hi[1]-lo[1], hi[2]-lo[2] }; if (d[0] < 0) return 0; return 2*(d[0]*d[1] + d[1]*d[2] + d[2]*d[0]); }
};

struct BvhBuilder {
  const std::vector<Box>& boxes;
  std::vector<int> prims;
  std::vector<BvhNode> nodes;
  explicit BvhBuilder(const std::vector<Box>& b) : boxes(b), prims(b.size()) { std::iota(prims.begin(), prims.end(), 0); }

  static float down(double v) { float f = (float)v; return f > v ? std::nextafter(f, -INFINITY) : f; }
  static float up(double v) { float f = (float)v; return f < v ? std::nextafter(f, INFINITY) : f; }

  void set_box(BvhNode& n, const Box& b) { for (int i = 0; i < 3; ++i) { n.lo[i] = down(b.lo[i]); n.hi[i] = up(b.hi[i]); } }

  void build() {
    nodes.reserve(2*boxes.size() + 2);
    nodes.push_back(BvhNode{});
    recurse(0, 0, (int)prims.size(), 0);
  }

  // Inner nodes in device layout: both child boxes in the parent (see BvhNode2).
  std::vector<BvhNode2> wide() const {
    // inner nodes are numbered breadth-first: the first k device nodes are the top of the tree, which is what the wavefront
    // traversal stages in shared memory when the whole tree does not fit
    std::vector<int> index(nodes.size(), -1);
    int n_inner = 0;
    if (!nodes.empty() && nodes[0].count == 0) {
      std::vector<int> queue(1, 0);
      for (size_t h = 0; h < queue.size(); ++h) {
        const int i = queue[h];
        index[(size_t)i] = n_inner++;
        for (int k = 0; k < 2; ++k) if (nodes[(size_t)nodes[(size_t)i].left + k].count == 0) queue.push_back(nodes[(size_t)i].left + k);
      }
    }
    std::vector<BvhNode2> out((size_t)std::max(n_inner, 1));
    auto set_child = [&](BvhNode2& w, int k, const BvhNode* c) {
      float* lo = k ? w.lo1 : w.lo0; float* hi = k ? w.hi1 : w.hi0;
      if (!c) { for (int a = 0; a < 3; ++a) { lo[a] = 3.0e38f; hi[a] = 3.0e38f; } w.child[k] = 0; w.count[k] = -1; return; }
      for (int a = 0; a < 3; ++a) { lo[a] = c->lo[a]; hi[a] = c->hi[a]; }
      if (c->count == 0) { w.child[k] = index[(size_t)(c - nodes.data())]; w.count[k] = 0; }
      else { w.child[k] = c->left; w.count[k] = c->count; }
    };
    if (n_inner == 0) {                     // the root is a leaf (or the scene is empty)
      set_child(out[0], 0, prims.empty() ? nullptr : &nodes[0]);
      set_child(out[0], 1, nullptr);
      return out;
    }
    for (size_t i = 0; i < nodes.size(); ++i) {
      if (nodes[i].count != 0) continue;
      BvhNode2& w = out[(size_t)index[i]];
      set_child(w, 0, &nodes[(size_t)nodes[i].left]);
      set_child(w, 1, &nodes[(size_t)nodes[i].left + 1]);
    }
    return out;
  }

  // 4-wide tree (Bvh4Node): every inner node of the binary tree that becomes a 4-wide node adopts up to four descendants —
  // its two children, the larger (by box area) of which is replaced by ITS children while there is room.  Nodes are numbered
  // breadth-first (the top of the tree first).  compact_of[face] >= 0: the face has a compact record (a whole sphere) and
  // its leaf entry is ~compact_of[face].  depth_out = levels of the 4-wide tree (bounds the traversal stack: 3 per level).
  void wide4(const std::vector<int>& compact_of, std::vector<Bvh4Node>& out, std::vector<int32_t>& prims4, int* depth_out) const {
    out.clear(); prims4.clear(); *depth_out = 1;
    auto area = [&](int i) { const BvhNode& n = nodes[(size_t)i];
                             const double d[3] = { (double)n.hi[0]-n.lo[0], (double)n.hi[1]-n.lo[1], (double)n.hi[2]-n.lo[2] };
                             return d[0]*d[1] + d[1]*d[2] + d[2]*d[0]; };
    auto leaf_ref = [&](const BvhNode& c) {
      const int first = (int)prims4.size();
      for (int k = 0; k < c.count; ++k) { const int f = prims[(size_t)c.left + k]; prims4.push_back(compact_of[(size_t)f] >= 0 ? ~compact_of[(size_t)f] : f); }
      return -2 - ((first << 3) | (std::min(c.count, 8) - 1));      // -1 is the traversal's "done"
    };
    auto empty_node = [] { Bvh4Node w; memset(&w, 0, sizeof w);
                           for (int a = 0; a < 3; ++a) for (int k = 0; k < 4; ++k) { w.b[2*a][k] = 3.0e38f; w.b[2*a+1][k] = -3.0e38f; }
                           for (int k = 0; k < 4; ++k) w.ref[k] = -1; return w; };
    auto set_slot = [&](Bvh4Node& w, int k, const BvhNode& c, int ref) {
      for (int a = 0; a < 3; ++a) { w.b[2*a][k] = c.lo[a]; w.b[2*a+1][k] = c.hi[a]; }
      w.ref[k] = ref;
    };
    if (prims.empty()) { out.push_back(empty_node()); return; }
    if (nodes[0].count != 0) {                        // the root is a leaf
      Bvh4Node w = empty_node(); set_slot(w, 0, nodes[0], leaf_ref(nodes[0])); out.push_back(w); return;
    }
    struct Pending { int bnode, level; };
    std::vector<Pending> queue(1, Pending{0, 1});
    std::vector<std::vector<int>> kids;               // binary node ids adopted by each 4-wide node
    for (size_t h = 0; h < queue.size(); ++h) {
      std::vector<int> c = { nodes[(size_t)queue[h].bnode].left, nodes[(size_t)queue[h].bnode].left + 1 };
      while (c.size() < 4) {
        int pick = -1;
        for (size_t k = 0; k < c.size(); ++k) if (nodes[(size_t)c[k]].count == 0 && (pick < 0 || area(c[k]) > area(c[(size_t)pick]))) pick = (int)k;
        if (pick < 0) break;
        const int b = c[(size_t)pick];
        c[(size_t)pick] = nodes[(size_t)b].left; c.push_back(nodes[(size_t)b].left + 1);
      }
      for (int b : c) if (nodes[(size_t)b].count == 0) queue.push_back(Pending{b, queue[h].level + 1});
      *depth_out = std::max(*depth_out, queue[h].level);
      kids.push_back(c);
    }
    // queue order = node numbering: the inner children of node h were appended in the order of kids[h]
    out.assign(queue.size(), empty_node());
    size_t next = 1;
    for (size_t h = 0; h < queue.size(); ++h)
      for (size_t k = 0; k < kids[h].size(); ++k) {
        const BvhNode& c = nodes[(size_t)kids[h][k]];
        if (c.count == 0) set_slot(out[h], (int)k, c, (int)next++); else set_slot(out[h], (int)k, c, leaf_ref(c));
      }
  }

  void recurse(int node, int first, int count, int depth) {
    Box bb; bb.reset(); Box cb; cb.reset();
    for (int i = first; i < first + count; ++i) {
      const Box& b = boxes[prims[i]]; bb.grow(b);
      for (int a = 0; a < 3; ++a) { double c = 0.5*(b.lo[a] + b.hi[a]); cb.lo[a] = std::min(cb.lo[a], c); cb.hi[a] = std::max(cb.hi[a], c); }
    }
    set_box(nodes[node], bb);
    const int LEAF = 1, NB = 16;            // one primitive per leaf: a leaf's box is the primitive's own box (tightest cull before the exact test)
    int best_axis = -1, best_split = -1; double best_cost = 1e300;
    if (count > LEAF && depth < 24) {      // deeper than 24: median splits only, so the depth stays below 24 + log2(n) < ODW_BVH_STACK
      for (int a = 0; a < 3; ++a) {
        double ext = cb.hi[a] - cb.lo[a];
        if (ext <= 0) continue;
        Box bins[NB]; int cnt[NB] = {0};
        for (auto& b : bins) b.reset();
        for (int i = first; i < first + count; ++i) {
          const Box& b = boxes[prims[i]];
          int k = std::min(NB-1, (int)(NB*((0.5*(b.lo[a] + b.hi[a]) - cb.lo[a])/ext)));
          bins[k].grow(b); cnt[k]++;
        }
        double right_area[NB]; int right_cnt[NB]; Box r; r.reset(); int rc = 0;
        for (int k = NB-1; k > 0; --k) { r.grow(bins[k]); rc += cnt[k]; right_area[k] = r.area(); right_cnt[k] = rc; }
        Box l; l.reset(); int lc = 0;
        for (int k = 0; k < NB-1; ++k) {
          l.grow(bins[k]); lc += cnt[k];
          if (lc == 0 || right_cnt[k+1] == 0) continue;
          double cost = l.area()*lc + right_area[k+1]*right_cnt[k+1];
          if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = k; }
        }
      }
    }
    if (best_axis < 0) {
      if (count <= (depth < 24 ? 8 : LEAF)) { nodes[node].left = first; nodes[node].count = count; return; }
      // degenerate (all centroids equal) or too deep: split in the middle
      int mid = first + count/2;
      int l = (int)nodes.size(); nodes.push_back(BvhNode{}); nodes.push_back(BvhNode{});
      nodes[node].left = l; nodes[node].count = 0;
      recurse(l, first, mid - first, depth + 1); recurse(l + 1, mid, first + count - mid, depth + 1);
      return;
    }
    double ext = cb.hi[best_axis] - cb.lo[best_axis];
    auto mid_it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](int pi) {
      const Box& b = boxes[pi];
      int k = std::min(NB-1, (int)(NB*((0.5*(b.lo[best_axis] + b.hi[best_axis]) - cb.lo[best_axis])/ext)));
      return k <= best_split;
    });
    int mid = (int)(mid_it - prims.begin());
    int l = (int)nodes.size(); nodes.push_back(BvhNode{}); nodes.push_back(BvhNode{});
    nodes[node].left = l; nodes[node].count = 0;
    recurse(l, first, mid - first, depth + 1); recurse(l + 1, mid, first + count - mid, depth + 1);
  }
};
}  // namespace

static std::vector<uint32_t> build_guide(const double* cdf, int n, int G = ODW_GUIDE);

// ---- convex shells ---------------------------------------------------------------------------------------------
// A shell bounds a convex solid iff every tangent plane of its surface is a supporting plane.  Checked on samples:
// corner / boundary points of planar faces, a 17x17 (u, v) grid (end points included) of sphere / cylinder / cone faces
// trimmed to a (u, v) box.  Anything else (torus, curved faces with loop trims, open shells of one face) is "not convex".
// Used to skip a shell for the segment that starts on it and points away from it (odw_trace.cuh interact()).
namespace {
struct Sample { double p[3], n[3]; };

void face_point_normal(const odw_face& f, double u, double v, Sample& s) {
  const double cu = std::cos(u), su = std::sin(u);
  double rad[3], g[3];
  for (int i = 0; i < 3; ++i) rad[i] = cu*f.xdir[i] + su*f.ydir[i];
  switch (f.kind) {
    case ODW_SURF_PLANE:
      for (int i = 0; i < 3; ++i) { s.p[i] = f.origin[i] + u*f.xdir[i] + v*f.ydir[i]; g[i] = f.zdir[i]; }
      break;
    case ODW_SURF_CYLINDER:
      for (int i = 0; i < 3; ++i) { s.p[i] = f.origin[i] + f.p0*rad[i] + v*f.zdir[i]; g[i] = rad[i]; }
      break;
    case ODW_SURF_CONE: {
      const double sa = std::sin(f.p1), ca = std::cos(f.p1), r = f.p0 + v*sa, sg = r >= 0 ? 1.0 : -1.0;
      for (int i = 0; i < 3; ++i) { s.p[i] = f.origin[i] + r*rad[i] + v*ca*f.zdir[i]; g[i] = sg*(ca*rad[i]*sg - sa*f.zdir[i]); }
      break;
    }
    default: {   // sphere
      const double cv = std::cos(v), sv = std::sin(v);
      for (int i = 0; i < 3; ++i) { g[i] = cv*rad[i] + sv*f.zdir[i]; s.p[i] = f.origin[i] + f.p0*g[i]; }
    }
  }
  const double l = std::sqrt(g[0]*g[0] + g[1]*g[1] + g[2]*g[2]);
  for (int i = 0; i < 3; ++i) s.n[i] = (double)f.nsign*g[i]/l;
}

bool shell_is_convex(const odw_scene_desc* sd, int face_first, int face_count) {
  if (face_count < 2 && !(face_count == 1 && sd->faces[face_first].kind == ODW_SURF_SPHERE && sd->faces[face_first].trim_kind == ODW_TRIM_NONE))
    return false;                                        // a single open face bounds nothing (a whole sphere does)
  std::vector<Sample> samples;
  double scale = 0;
  for (int fi = face_first; fi < face_first + face_count; ++fi) {
    const odw_face& f = sd->faces[fi];
    if (f.kind == ODW_SURF_TORUS || f.kind == ODW_SURF_CONICOID) return false;
    Sample s;
    if (f.kind == ODW_SURF_PLANE) {
      if (f.trim_kind == ODW_TRIM_UVBOX) {
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) { face_point_normal(f, a ? f.uv_max[0] : f.uv_min[0], b ? f.uv_max[1] : f.uv_min[1], s); samples.push_back(s); }
      } else if (f.trim_kind == ODW_TRIM_LOOPS) {
        for (int k = f.seg_first; k < f.seg_first + f.seg_count; ++k) {
          const odw_trimseg& g = sd->segs[k];
          if (g.kind == ODW_SEG_LINE) { face_point_normal(f, g.a[0], g.a[1], s); samples.push_back(s); face_point_normal(f, g.a[2], g.a[3], s); samples.push_back(s); }
          else for (int m = 0; m <= 16; ++m) { const double ang = g.a[3] + g.a[4]*m/16.0; face_point_normal(f, g.a[0] + g.a[2]*std::cos(ang), g.a[1] + g.a[2]*std::sin(ang), s); samples.push_back(s); }
        }
      } else return false;
    } else {
      if (f.trim_kind == ODW_TRIM_LOOPS) return false;
      double u0 = f.uv_min[0], u1 = f.uv_max[0], v0 = f.uv_min[1], v1 = f.uv_max[1];
      if (f.trim_kind == ODW_TRIM_NONE) { u0 = 0; u1 = ODW_TWO_PI; v0 = -ODW_TWO_PI/4; v1 = ODW_TWO_PI/4; }
      for (int a = 0; a <= 16; ++a) for (int b = 0; b <= 16; ++b) { face_point_normal(f, u0 + (u1 - u0)*a/16.0, v0 + (v1 - v0)*b/16.0, s); samples.push_back(s); }
    }
    if (samples.size() > 20000) return false;            // huge shells (tessellated bodies): not worth the quadratic test
  }
  for (const Sample& s : samples) for (int i = 0; i < 3; ++i) scale = std::max(scale, std::fabs(s.p[i]));
  const double eps = 1e-9*std::max(1.0, scale);
  for (const Sample& a : samples)
    for (const Sample& b : samples) {
      const double d = a.n[0]*(b.p[0] - a.p[0]) + a.n[1]*(b.p[1] - a.p[1]) + a.n[2]*(b.p[2] - a.p[2]);
      if (d > eps) return false;
    }
  return true;
}
}  // namespace

// odw_face -> the record the kernels read.  fast_paths: precompute the inline-test constants of the trace kernel
// (emitting faces of a surface source are only evaluated / trim-tested, never intersected: no fast paths there).
static void fill_dface(const odw_face& f, const odw_trimseg* segs, bool fast_paths, DFace& d) {
    memset(&d, 0, sizeof d);
    for (int k = 0; k < 3; ++k) { d.o[k] = f.origin[k]; d.x[k] = f.xdir[k]; d.y[k] = f.ydir[k]; d.z[k] = f.zdir[k]; }
    d.p0 = f.p0; d.p1 = f.p1;
    d.umin = f.uv_min[0]; d.umax = f.uv_max[0]; d.vmin = f.uv_min[1]; d.vmax = f.uv_max[1];
    d.kind = f.kind; d.trim = f.trim_kind; d.nsign = f.nsign; d.group = f.group;
    d.seg_first = f.seg_first; d.seg_count = f.seg_count; d.face_id = f.face_id;
    if (f.kind == ODW_SURF_CONICOID && f.seg_count > 0 && segs && segs[f.seg_first].kind == ODW_SEG_ASPHERE) {
      // the auxiliary record moves into the face record; the kernels see only boundary pieces in the segment range
      for (int k = 0; k < 5; ++k) d.aux[k] = segs[f.seg_first].a[k];
      d.aux[5] = 1.0;
      d.seg_first = f.seg_first + 1; d.seg_count = f.seg_count - 1;
    }
    d.flags = (f.kind != ODW_SURF_PLANE && std::fabs((f.uv_max[0] - f.uv_min[0]) - ODW_TWO_PI) < 1e-9) ? DFACE_FULL_U : 0;
    {
      auto dot = [](const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; };
      const bool full_u = (d.flags & DFACE_FULL_U) != 0;
      if (fast_paths) {
      if (f.kind == ODW_SURF_PLANE && f.trim_kind == ODW_TRIM_UVBOX) {
        d.flags |= DFACE_FAST; d.c0 = dot(d.o, d.z); d.c1 = dot(d.o, d.x); d.c2 = dot(d.o, d.y);
      } else if (f.kind == ODW_SURF_PLANE && f.trim_kind == ODW_TRIM_LOOPS && f.seg_count == 1 &&
                 segs[f.seg_first].kind == ODW_SEG_ARC && segs[f.seg_first].a[4] >= ODW_TWO_PI - 1e-12) {
        // a disc: the single trim loop is one full circle (centre a[0], a[1], radius a[2])
        const odw_trimseg& c = segs[f.seg_first];
        d.flags |= DFACE_FAST | DFACE_DISC; d.c0 = dot(d.o, d.z); d.c1 = dot(d.o, d.x); d.c2 = dot(d.o, d.y);
        d.umin = c.a[0]; d.vmin = c.a[1]; d.umax = c.a[2]; d.vmax = c.a[2];
      } else if (f.kind == ODW_SURF_SPHERE && (f.trim_kind == ODW_TRIM_NONE || (f.trim_kind == ODW_TRIM_UVBOX && full_u))) {
        // axial window [c0, c1] of the zone; the tolerance is a distance ALONG the sphere (ray.py:426 measures the distance to
        // the trimmed face), i.e. tol*cos(v) in the axial coordinate: aux[0], aux[1] carry the two cosines.  A bound at a pole
        // is no bound (and must not reject a hit whose axial coordinate exceeds R by a rounding error).
        d.flags |= DFACE_FAST;
        const bool whole = f.trim_kind == ODW_TRIM_NONE;
        const bool lo_open = whole || f.uv_min[1] <= -ODW_TWO_PI/4 + 1e-12, hi_open = whole || f.uv_max[1] >= ODW_TWO_PI/4 - 1e-12;
        d.c0 = lo_open ? -1e300 : f.p0*std::sin(f.uv_min[1]);
        d.c1 = hi_open ?  1e300 : f.p0*std::sin(f.uv_max[1]);
        d.aux[0] = lo_open ? 0.0 : std::cos(f.uv_min[1]);
        d.aux[1] = hi_open ? 0.0 : std::cos(f.uv_max[1]);
      } else if (f.kind == ODW_SURF_CYLINDER && f.trim_kind == ODW_TRIM_UVBOX && full_u) {
        d.flags |= DFACE_FAST; d.c0 = f.uv_min[1]; d.c1 = f.uv_max[1]; d.aux[0] = d.aux[1] = 1.0;
      }
      }
    }
}

extern "C" int odw_scene_create(odw_engine* eng, const odw_scene_desc* sd, odw_scene** out) {
  if (!eng || !sd || !out) return fail(ODW_EINVAL, "odw_scene_create: NULL argument");
  *out = nullptr;
  if (sd->n_faces < 0 || sd->n_groups <= 0 || (sd->n_faces > 0 && !sd->faces) || !sd->groups)
    return fail(ODW_EINVAL, "odw_scene_create: empty or inconsistent description");
  if (sd->n_seq_steps > 128) return fail(ODW_EUNSUPPORTED, "more than 128 sequential steps");
  CU(cudaSetDevice(eng->device));
  std::vector<DFace> faces((size_t)sd->n_faces);
  std::vector<Box> boxes((size_t)sd->n_faces);
  for (int i = 0; i < sd->n_faces; ++i) {
    const odw_face& f = sd->faces[i];
    if (f.group < 0 || f.group >= sd->n_groups) return fail(ODW_EINVAL, "face " + std::to_string(i) + ": group out of range");
    if (f.kind < ODW_SURF_PLANE || f.kind > ODW_SURF_CONICOID) return fail(ODW_EINVAL, "face " + std::to_string(i) + ": unknown surface kind");
    if (f.kind == ODW_SURF_CONICOID && !(f.p0 != 0 && std::isfinite(f.p0) && std::isfinite(f.p1)))
      return fail(ODW_EINVAL, "face " + std::to_string(i) + ": a conicoid needs a finite non-zero vertex curvature (p0) and a finite conic constant (p1)");
    if ((f.trim_kind == ODW_TRIM_LOOPS || (f.kind == ODW_SURF_CONICOID && f.seg_count > 0)) &&
        (f.seg_first < 0 || f.seg_count < 0 || f.seg_first + f.seg_count > sd->n_segs))
      return fail(ODW_EINVAL, "face " + std::to_string(i) + ": trim segment range out of bounds");
    for (int k = 0; k < ((f.trim_kind == ODW_TRIM_LOOPS || f.kind == ODW_SURF_CONICOID) ? f.seg_count : 0); ++k)
      if (sd->segs[f.seg_first + k].kind == ODW_SEG_ASPHERE && (k != 0 || f.kind != ODW_SURF_CONICOID))
        return fail(ODW_EINVAL, "face " + std::to_string(i) + ": an asphere record must be the first segment of a conicoid face");
    DFace& d = faces[(size_t)i];
    fill_dface(f, sd->segs, true, d);
    for (int k = 0; k < 3; ++k) { boxes[(size_t)i].lo[k] = f.aabb_min[k]; boxes[(size_t)i].hi[k] = f.aabb_max[k]; }
    for (int s = 0; s < sd->n_seq_steps; ++s)
      for (int k = sd->seq_offsets[s]; k < sd->seq_offsets[s+1]; ++k)
        if (sd->seq_groups[k] == f.group) d.seqmask[s >> 6] |= 1ull << (s & 63);
  }
  // shells: the reference culls per shell box first; faces of one shell must be contiguous (odw.h: "sorted by shell")
  std::vector<DShell> shells;
  double extent = 0;                           // max |coordinate| of the scene (feeds the fp32 culling margin)
  bool skip_convex = true;
  if (const char* w = getenv("ODW_SKIP_CONVEX")) skip_convex = atoi(w) != 0;
  auto make_shell = [&](int face_first, int face_count, int group) {
    DShell d; memset(&d, 0, sizeof d);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int f = face_first; f < face_first + face_count; ++f)
      for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], sd->faces[f].aabb_min[k]); hi[k] = std::max(hi[k], sd->faces[f].aabb_max[k]); }
    for (int k = 0; k < 3; ++k) {
      d.lo[k] = BvhBuilder::down(lo[k]); d.hi[k] = BvhBuilder::up(hi[k]);
      if (face_count > 0) extent = std::max(extent, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
    }
    d.face_first = face_first; d.face_count = face_count; d.group = group;
    d.convex = (skip_convex && shell_is_convex(sd, face_first, face_count)) ? 1 : 0;
    for (int f = face_first; f < face_first + face_count; ++f) faces[(size_t)f].shell = (int32_t)shells.size();
    if (face_count > 0) { d.seqmask[0] = faces[(size_t)face_first].seqmask[0]; d.seqmask[1] = faces[(size_t)face_first].seqmask[1]; }
    return d;
  };
  auto pair_planes = [&](int face_first, int face_count) {
    // opposite faces of a box come as consecutive planes with the SAME normal vector (the orientation is in nsign): the first
    // of such a pair is flagged, the shared-memory kernel then tests both with one reciprocal and one set of dot products
    for (int f = face_first; f + 1 < face_first + face_count; ++f) {
      DFace& a = faces[(size_t)f]; const DFace& b = faces[(size_t)f + 1];
      const bool planes = a.kind == ODW_SURF_PLANE && b.kind == ODW_SURF_PLANE && (a.flags & DFACE_FAST) && (b.flags & DFACE_FAST);
      if (planes && a.z[0] == b.z[0] && a.z[1] == b.z[1] && a.z[2] == b.z[2]) { a.flags |= DFACE_PAIR; ++f; }
    }
  };
  if (sd->n_shells > 0 && sd->shells) {
    for (int i = 0; i < sd->n_shells; ++i) {
      const odw_shell& h = sd->shells[i];
      if (h.face_first < 0 || h.face_count < 0 || h.face_first + h.face_count > sd->n_faces || h.group < 0 || h.group >= sd->n_groups)
        return fail(ODW_EINVAL, "shell " + std::to_string(i) + ": face range or group out of bounds");
      for (int f = h.face_first; f < h.face_first + h.face_count; ++f)
        if (sd->faces[f].group != h.group) return fail(ODW_EINVAL, "shell " + std::to_string(i) + ": faces of a shell must share its group");
      shells.push_back(make_shell(h.face_first, h.face_count, h.group));
      pair_planes(h.face_first, h.face_count);
    }
    std::vector<char> covered((size_t)sd->n_faces, 0);
    for (const DShell& d : shells) for (int f = d.face_first; f < d.face_first + d.face_count; ++f) covered[(size_t)f]++;
    for (int f = 0; f < sd->n_faces; ++f) if (covered[(size_t)f] != 1) return fail(ODW_EINVAL, "face " + std::to_string(f) + " is not covered by exactly one shell");
  } else {
    // no shell table given: one shell per face
    for (int f = 0; f < sd->n_faces; ++f) shells.push_back(make_shell(f, 1, sd->faces[f].group));
  }
  std::vector<DGroup> groups((size_t)sd->n_groups);
  for (int i = 0; i < sd->n_groups; ++i) {
    const odw_group& g = sd->groups[i]; DGroup& d = groups[(size_t)i];
    d.n = g.refractive_index; d.reflectivity = g.reflectivity; d.absorption_length = g.absorption_length;
    d.lpm = g.grating_lines_per_mm; d.order = g.grating_order;
    for (int k = 0; k < 3; ++k) d.gdir[k] = g.grating_orientation[k];
    d.type = g.optical_type; d.record = g.record_hits; d.gtype = g.grating_type;
    d.fresnel = (g.fresnel && g.optical_type == ODW_OPT_LENS) ? 1 : 0;
    d.scat_main = d.scat_modify = -1; d.pad1 = d.pad2 = 0;
    if (sd->n_scatters > 0 && sd->group_scatter && (g.optical_type == ODW_OPT_MIRROR || g.optical_type == ODW_OPT_LENS)) {
      d.scat_main = sd->group_scatter[2*i]; d.scat_modify = sd->group_scatter[2*i+1];
      if (d.scat_main >= sd->n_scatters || d.scat_modify >= sd->n_scatters) return fail(ODW_EINVAL, "group " + std::to_string(i) + ": scatter index out of range");
    }
  }
  for (int i = 0; i < sd->n_scatters; ++i) {
    const odw_scatter& t = sd->scatters[i];
    if (t.n_first < 2 || t.n_phi < 2 || !t.phi_cdf || !t.first_cdf || (t.n_rows != 1 && t.n_rows != t.n_phi - 1))
      return fail(ODW_EINVAL, "scatter " + std::to_string(i) + ": malformed tables");
    if (t.n_tables > 1 && t.n_rows != 1) return fail(ODW_EUNSUPPORTED, "scatter " + std::to_string(i) + ": a per-hit family of tables needs n_rows == 1");
  }
  for (int i = 0; i < sd->n_groups && sd->n_scatters > 0 && sd->group_scatter; ++i) {
    const int m = sd->group_scatter[2*i], mod = sd->group_scatter[2*i+1];
    if (m >= 0 && m < sd->n_scatters && sd->scatters[m].n_tables > 1 && sd->groups[i].optical_type == ODW_OPT_LENS && (sd->scatters[m].n_tables & 1))
      return fail(ODW_EINVAL, "group " + std::to_string(i) + ": the per-hit family of a Lens group holds two families (entering, leaving): n_tables must be even");
    if (mod >= 0 && mod < sd->n_scatters && sd->scatters[mod].n_tables > 1)
      return fail(ODW_EINVAL, "group " + std::to_string(i) + ": the ray-modification density is never re-compiled per hit (optical_group.py:318): one table");
  }
  odw_scene* sc = new odw_scene();
  sc->eng = eng; sc->n_groups = sd->n_groups; sc->extent = extent;
  for (const DFace& f : faces) if (f.aux[5] != 0.0) sc->ext_optics = true;      // even-asphere faces need the FEAT_EXT instances
  for (const DGroup& g : groups)
    if (g.type == ODW_OPT_GRATING || g.scat_main >= 0 || g.scat_modify >= 0 || std::isfinite(g.absorption_length) || g.fresnel) sc->ext_optics = true;
  int rc;
  if ((rc = upload(eng, sc->owned, faces.data(), faces.size(), &sc->d.faces))) { odw_scene_destroy(sc); return rc; }
  if ((rc = upload(eng, sc->owned, shells.data(), shells.size(), &sc->d.shells))) { odw_scene_destroy(sc); return rc; }
  sc->d.n_shells = (int)shells.size();
  if ((rc = upload(eng, sc->owned, sd->segs, (size_t)sd->n_segs, &sc->d.segs))) { odw_scene_destroy(sc); return rc; }
  if ((rc = upload(eng, sc->owned, groups.data(), groups.size(), &sc->d.groups))) { odw_scene_destroy(sc); return rc; }
  if (sd->n_scatters > 0) {
    std::vector<DScatter> scat((size_t)sd->n_scatters);
    for (int i = 0; i < sd->n_scatters; ++i) {
      const odw_scatter& t = sd->scatters[i]; DScatter& d = scat[(size_t)i];
      const size_t nt = (size_t)std::max(1, t.n_tables);                          // members of a per-hit family, one after the other
      if ((rc = upload(eng, sc->owned, t.phi_cdf, nt*(size_t)t.n_phi, &d.phi_cdf))) { odw_scene_destroy(sc); return rc; }
      if ((rc = upload(eng, sc->owned, t.first_cdf, nt*(size_t)t.n_rows*t.n_first, &d.first_cdf))) { odw_scene_destroy(sc); return rc; }
      std::vector<uint32_t> pg, fg;
      for (size_t m = 0; m < nt; ++m) { std::vector<uint32_t> g = build_guide(t.phi_cdf + m*(size_t)t.n_phi, t.n_phi); pg.insert(pg.end(), g.begin(), g.end()); }
      for (size_t r = 0; r < nt*(size_t)t.n_rows; ++r) { std::vector<uint32_t> g = build_guide(t.first_cdf + r*(size_t)t.n_first, t.n_first); fg.insert(fg.end(), g.begin(), g.end()); }
      if ((rc = upload(eng, sc->owned, pg.data(), pg.size(), &d.phi_guide))) { odw_scene_destroy(sc); return rc; }
      if ((rc = upload(eng, sc->owned, fg.data(), fg.size(), &d.first_guide))) { odw_scene_destroy(sc); return rc; }
      d.first_lo = t.first_lo; d.first_hi = t.first_hi; d.phi_lo = t.phi_lo; d.phi_hi = t.phi_hi;
      d.n_first = t.n_first; d.n_phi = t.n_phi; d.n_rows = t.n_rows; d.n_tables = std::max(1, t.n_tables);
    }
    if ((rc = upload(eng, sc->owned, scat.data(), scat.size(), &sc->d.scatters))) { odw_scene_destroy(sc); return rc; }
  }
  sc->d.n_faces = sd->n_faces; sc->d.n_segs = sd->n_segs; sc->d.n_groups = sd->n_groups; sc->d.n_seq_steps = sd->n_seq_steps;
  sc->use_bvh = sd->n_faces > SMEM_FACE_LIMIT;
  if (const char* w = getenv("ODW_BVH")) { if (atoi(w) == 1 && sd->n_faces > 0) sc->use_bvh = true; }   // developer/test knob: BVH path for small scenes too
  if (const char* w = getenv("ODW_WAVEFRONT")) sc->wavefront = atoi(w) != 0;
  if (sc->use_bvh) {
    BvhBuilder b(boxes);
    b.build();
    sc->bvh_host = b.wide();
    const BvhNode2* dev = nullptr;
    if ((rc = upload(eng, sc->owned, sc->bvh_host.data(), sc->bvh_host.size(), &dev))) { odw_scene_destroy(sc); return rc; }
    sc->bvh_dev = const_cast<BvhNode2*>(dev); sc->d.bvh = dev; sc->bvh_margin = 0.0f;
    if ((rc = upload(eng, sc->owned, b.prims.data(), b.prims.size(), &sc->d.bvh_prims))) { odw_scene_destroy(sc); return rc; }
    sc->d.n_bvh_nodes = (int)sc->bvh_host.size();
    sc->smem = 0;
    // 4-wide tree + compact sphere records for the wavefront traversal (odw_wavefront.cu wf_traverse4)
    {
      std::vector<int> compact_of((size_t)sd->n_faces, -1);
      std::vector<DSphere> spheres; std::vector<int2> info;
      for (int i = 0; i < sd->n_faces; ++i) {
        const odw_face& f = sd->faces[i];
        if (f.kind == ODW_SURF_SPHERE && f.trim_kind == ODW_TRIM_NONE) {
          compact_of[(size_t)i] = (int)spheres.size();
          spheres.push_back(DSphere{ f.origin[0], f.origin[1], f.origin[2], f.p0 });
          info.push_back(make_int2(i, f.group));
        }
      }
      std::vector<int32_t> prims4; int depth4 = 1;
      b.wide4(compact_of, sc->bvh4_host, prims4, &depth4);
      bool use4 = 3*depth4 + 1 <= ODW_BVH_STACK;                    // the traversal stacks up to three siblings per level
      if (const char* w = getenv("ODW_BVH4")) use4 = use4 && atoi(w) != 0;
      if (use4) {
        const Bvh4Node* dev4 = nullptr;
        if ((rc = upload(eng, sc->owned, sc->bvh4_host.data(), sc->bvh4_host.size(), &dev4))) { odw_scene_destroy(sc); return rc; }
        sc->bvh4_dev = const_cast<Bvh4Node*>(dev4); sc->d.bvh4 = dev4; sc->d.n_bvh4_nodes = (int)sc->bvh4_host.size();
        if ((rc = upload(eng, sc->owned, prims4.data(), prims4.size(), &sc->d.bvh4_prims))) { odw_scene_destroy(sc); return rc; }
        sc->d.n_bvh4_prims = (int)prims4.size();
        if ((rc = upload(eng, sc->owned, spheres.data(), spheres.size(), &sc->d.spheres))) { odw_scene_destroy(sc); return rc; }
        if ((rc = upload(eng, sc->owned, info.data(), info.size(), &sc->d.sphere_info))) { odw_scene_destroy(sc); return rc; }
        sc->d.n_spheres = (int)spheres.size(); sc->d.bvh_depth = depth4;
        std::vector<ulonglong2> gmask((size_t)sd->n_groups, make_ulonglong2(0ull, 0ull));
        for (int st = 0; st < sd->n_seq_steps; ++st)
          for (int k = sd->seq_offsets[st]; k < sd->seq_offsets[st+1]; ++k) {
            const int g = sd->seq_groups[k];
            if (g >= 0 && g < sd->n_groups) { if (st < 64) gmask[(size_t)g].x |= 1ull << st; else gmask[(size_t)g].y |= 1ull << (st - 64); }
          }
        if ((rc = upload(eng, sc->owned, gmask.data(), gmask.size(), &sc->d.group_seqmask))) { odw_scene_destroy(sc); return rc; }
      }
    }
  } else {
    sc->smem = std::max<size_t>(16, faces.size()*sizeof(DFace) + shells.size()*sizeof(DShell));
  }
  *out = sc;
  return ODW_OK;
}

extern "C" void odw_scene_destroy(odw_scene* sc) {
  if (!sc) return;
  cudaSetDevice(sc->eng->device);
  for (void* p : sc->owned) cudaFree(p);
  delete sc;
}

static std::vector<uint32_t> build_guide(const double* cdf, int n, int G) {
  std::vector<uint32_t> g((size_t)G + 1);
  for (int k = 0; k <= G; ++k) {
    double x = (double)k/(double)G;
    const double* it = std::upper_bound(cdf, cdf + n, x);     // first element > x
    long j = (it - cdf) - 1;
    g[(size_t)k] = (uint32_t)std::max<long>(0, j);
  }
  return g;
}

extern "C" void odw_source_destroy(odw_source* s);

// A plane face trimmed to exactly one triangle (three chained straight pcurves): what the tessellation of a free-form emitter
// consists of.  tri = (u, v) of the three corners.  Same rule as the oracle's emit_triangle.
static bool emit_triangle(const odw_face& f, const odw_trimseg* segs, double* tri) {
  if (f.kind != ODW_SURF_PLANE || f.trim_kind != ODW_TRIM_LOOPS || f.seg_count != 3 || !segs) return false;
  const odw_trimseg* g = segs + f.seg_first;
  double scale = 0;
  for (int k = 0; k < 3; ++k) { if (g[k].kind != ODW_SEG_LINE) return false; for (int j = 0; j < 4; ++j) scale = std::max(scale, std::fabs(g[k].a[j])); }
  const double eps = 1e-9*std::max(scale, 1e-300);
  for (int k = 0; k < 3; ++k) {
    const odw_trimseg& a = g[k]; const odw_trimseg& b = g[(k + 1) % 3];
    if (std::fabs(a.a[2] - b.a[0]) > eps || std::fabs(a.a[3] - b.a[1]) > eps) return false;
    tri[2*k] = a.a[0]; tri[2*k + 1] = a.a[1];
  }
  return true;
}

// ODW_SRC_SURFACE (reference freecad_elements/surface_source.py): emitting faces, area CDF and the theta table
static int surface_source_create(odw_engine* eng, const odw_source_desc* sd, odw_source** out) {
  if (sd->n_emit <= 0 || !sd->emit_faces || !sd->emit_cdf) return fail(ODW_EINVAL, "surface source without emitting faces");
  if (sd->n_first < 2 || !sd->first_cdf) return fail(ODW_EINVAL, "surface source: missing theta CDF");
  if (!(sd->first_cdf[0] == 0.0) || !(std::fabs(sd->first_cdf[sd->n_first-1] - 1.0) < 1e-12)) return fail(ODW_EINVAL, "surface source: theta CDF must run from 0 to 1");
  if (!(std::fabs(sd->emit_cdf[sd->n_emit-1] - 1.0) < 1e-12)) return fail(ODW_EINVAL, "surface source: emit_cdf must end at 1");
  for (int i = 0; i < sd->n_emit; ++i) {
    const odw_face& f = sd->emit_faces[i];
    if (f.kind < ODW_SURF_PLANE || f.kind > ODW_SURF_TORUS) return fail(ODW_EINVAL, "surface source: unknown surface kind of an emitting face");   // conicoids do not emit (no area-uniform draw yet): mesh them
    if (f.trim_kind == ODW_TRIM_LOOPS && (f.seg_first < 0 || f.seg_first + f.seg_count > sd->n_emit_segs || !sd->emit_segs))
      return fail(ODW_EINVAL, "surface source: trim segment range of an emitting face out of bounds");
    if (i && sd->emit_cdf[i] < sd->emit_cdf[i-1]) return fail(ODW_EINVAL, "surface source: emit_cdf must be non-decreasing");
  }
  CU(cudaSetDevice(eng->device));
  odw_source* s = new odw_source();
  s->eng = eng;
  std::vector<DFace> faces((size_t)sd->n_emit);
  double bound = 0;
  for (int i = 0; i < sd->n_emit; ++i) {
    fill_dface(sd->emit_faces[i], sd->emit_segs, false, faces[(size_t)i]);
    double tri[6];
    if (emit_triangle(sd->emit_faces[i], sd->emit_segs, tri)) {      // tessellated emitters: sampled without rejection (odw_trace.cuh init_ray_surface)
      faces[(size_t)i].flags |= DFACE_TRI;
      for (int k = 0; k < 6; ++k) faces[(size_t)i].aux[k] = tri[k];
    }
    for (int k = 0; k < 3; ++k) bound = std::max(bound, std::max(std::fabs(sd->emit_faces[i].aabb_min[k]), std::fabs(sd->emit_faces[i].aabb_max[k])));
  }
  int rc;
  if ((rc = upload(eng, s->owned, faces.data(), faces.size(), &s->d.emit_faces))) { odw_source_destroy(s); return rc; }
  if ((rc = upload(eng, s->owned, sd->emit_segs, (size_t)sd->n_emit_segs, &s->d.emit_segs))) { odw_source_destroy(s); return rc; }
  if ((rc = upload(eng, s->owned, sd->emit_cdf, (size_t)sd->n_emit, &s->d.emit_cdf))) { odw_source_destroy(s); return rc; }
  {
    std::vector<uint32_t> eg((size_t)ODW_EMIT_GUIDE + 1);
    for (int k = 0; k <= ODW_EMIT_GUIDE; ++k) {                       // first face with emit_cdf > k/ODW_EMIT_GUIDE
      const double x = (double)k/(double)ODW_EMIT_GUIDE;
      const long j = std::upper_bound(sd->emit_cdf, sd->emit_cdf + sd->n_emit, x) - sd->emit_cdf;
      eg[(size_t)k] = (uint32_t)std::min<long>(j, sd->n_emit - 1);
    }
    if ((rc = upload(eng, s->owned, eg.data(), eg.size(), &s->d.emit_guide))) { odw_source_destroy(s); return rc; }
  }
  if ((rc = upload(eng, s->owned, sd->first_cdf, (size_t)sd->n_first, &s->d.first_cdf))) { odw_source_destroy(s); return rc; }
  int theta_cells = ODW_GUIDE;
  while (theta_cells < sd->n_first && theta_cells < (1 << 20)) theta_cells <<= 1;
  s->d.n_first_guide = theta_cells;
  std::vector<uint32_t> fg = build_guide(sd->first_cdf, sd->n_first, theta_cells);
  if ((rc = upload(eng, s->owned, fg.data(), fg.size(), &s->d.first_guide))) { odw_source_destroy(s); return rc; }
  s->d.first_lo = sd->first_lo; s->d.first_hi = sd->first_hi; s->d.phi_lo = 0; s->d.phi_hi = ODW_TWO_PI;
  s->d.wavelength = sd->wavelength;
  s->d.kind = sd->kind; s->d.source_id = sd->source_id; s->d.n_first = sd->n_first; s->d.n_phi = 0; s->d.n_rows = 1;
  s->d.n_emit = sd->n_emit; s->d.dist_tol = sd->dist_tol > 0 ? sd->dist_tol : 1e-6;
  s->origin_bound = bound;
  s->max_ray_length_scale = sd->max_ray_length_scale > 0 ? sd->max_ray_length_scale : 1.0;
  s->max_intersections_scale = sd->max_intersections_scale > 0 ? sd->max_intersections_scale : 1.0;
  if (sd->n_ignored > 0 && sd->ignored_groups) s->ignored.assign(sd->ignored_groups, sd->ignored_groups + sd->n_ignored);
  *out = s;
  return ODW_OK;
}

extern "C" int odw_source_create(odw_engine* eng, const odw_source_desc* sd, odw_source** out) {
  if (!eng || !sd || !out) return fail(ODW_EINVAL, "odw_source_create: NULL argument");
  *out = nullptr;
  if (sd->kind == ODW_SRC_SURFACE) return surface_source_create(eng, sd, out);
  if (sd->kind != ODW_SRC_POINT_SPHERICAL && sd->kind != ODW_SRC_POINT_COLLIMATED) return fail(ODW_EUNSUPPORTED, "unknown source kind");
  if (sd->n_first < 2 || sd->n_phi < 2 || !sd->phi_cdf || !sd->first_cdf) return fail(ODW_EINVAL, "odw_source_create: missing CDF tables");
  if (sd->n_rows != 1 && sd->n_rows != sd->n_phi - 1) return fail(ODW_EINVAL, "odw_source_create: n_rows must be 1 or n_phi-1");
  for (int r = 0; r < sd->n_rows + 1; ++r) {
    const double* c = r < sd->n_rows ? sd->first_cdf + (size_t)r*sd->n_first : sd->phi_cdf;
    int n = r < sd->n_rows ? sd->n_first : sd->n_phi;
    if (!(c[0] == 0.0) || !(std::fabs(c[n-1] - 1.0) < 1e-12)) return fail(ODW_EINVAL, "odw_source_create: CDF rows must run from 0 to 1");
  }
  CU(cudaSetDevice(eng->device));
  odw_source* s = new odw_source();
  s->eng = eng;
  int rc;
  if ((rc = upload(eng, s->owned, sd->phi_cdf, (size_t)sd->n_phi, &s->d.phi_cdf))) { odw_source_destroy(s); return rc; }
  if ((rc = upload(eng, s->owned, sd->first_cdf, (size_t)sd->n_rows*sd->n_first, &s->d.first_cdf))) { odw_source_destroy(s); return rc; }
  std::vector<uint32_t> pg = build_guide(sd->phi_cdf, sd->n_phi), fg;
  fg.reserve((size_t)sd->n_rows*(ODW_GUIDE + 1));
  for (int r = 0; r < sd->n_rows; ++r) {
    std::vector<uint32_t> g = build_guide(sd->first_cdf + (size_t)r*sd->n_first, sd->n_first);
    fg.insert(fg.end(), g.begin(), g.end());
  }
  if ((rc = upload(eng, s->owned, pg.data(), pg.size(), &s->d.phi_guide))) { odw_source_destroy(s); return rc; }
  if ((rc = upload(eng, s->owned, fg.data(), fg.size(), &s->d.first_guide))) { odw_source_destroy(s); return rc; }
  s->d.first_lo = sd->first_lo; s->d.first_hi = sd->first_hi; s->d.phi_lo = sd->phi_lo; s->d.phi_hi = sd->phi_hi;
  s->d.focal = sd->focal_length; s->d.wavelength = sd->wavelength;
  for (int i = 0; i < 12; ++i) s->d.M[i] = sd->gpM[i];
  s->d.kind = sd->kind; s->d.source_id = sd->source_id; s->d.n_first = sd->n_first; s->d.n_phi = sd->n_phi; s->d.n_rows = sd->n_rows;
  {
    // _makeRay: |local origin| <= 2|f| per component (spherical) or max |r| (collimated), then gpM
    const double r = sd->kind == ODW_SRC_POINT_SPHERICAL ? 2*std::fabs(sd->focal_length)
                                                         : std::max(std::fabs(sd->first_lo), std::fabs(sd->first_hi));
    double b = 0;
    for (int i = 0; i < 3; ++i) {
      double row = std::fabs(sd->gpM[4*i+3]);
      for (int k = 0; k < 3; ++k) row += r*std::fabs(sd->gpM[4*i+k]);
      b = std::max(b, row);
    }
    s->origin_bound = b;
  }
  s->max_ray_length_scale = sd->max_ray_length_scale > 0 ? sd->max_ray_length_scale : 1.0;
  s->max_intersections_scale = sd->max_intersections_scale > 0 ? sd->max_intersections_scale : 1.0;
  if (sd->n_ignored > 0 && sd->ignored_groups) s->ignored.assign(sd->ignored_groups, sd->ignored_groups + sd->n_ignored);
  *out = s;
  return ODW_OK;
}

extern "C" void odw_source_destroy(odw_source* s) {
  if (!s) return;
  cudaSetDevice(s->eng->device);
  for (void* p : s->owned) cudaFree(p);
  delete s;
}

// ------------------------------------------------------------------------------------------
extern "C" void odw_result_destroy(odw_result* r) {
  if (!r) return;
  odw_engine* e = r->eng;
  cudaSetDevice(e->device);
  void* ptrs[] = { r->hb.points, r->hb.dirs, r->hb.powers, r->hb.entering, r->hb.ray_index, r->hb.group, r->hb.bounce,
                   r->hb.face_id, r->hb.medium, r->dcounters, r->dbins, r->dbinnings, r->d_nseg, r->d_final_point, r->d_final_power, r->d_final_medium,
                   r->d_in_o, r->d_in_d, r->d_in_p };
  for (void* p : ptrs) e->release(p);
  delete r;
}

static int prepare_result(odw_engine* eng, const odw_scene* sc, const odw_trace_cfg* cfg, uint64_t n_rays, odw_result** out,
                          TraceParams& p) {
  odw_result* r = new odw_result();
  r->eng = eng; r->n_rays = n_rays;
  *out = r;
  int rc;
  uint64_t cap = cfg->store_hits ? (cfg->hit_capacity ? cfg->hit_capacity : std::max<uint64_t>(1024, 2*n_rays)) : 0;
  r->hb.capacity = cap;
  if (cap) {
    if ((rc = eng->alloc((void**)&r->hb.points, cap*24))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.dirs, cap*24))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.powers, cap*8))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.entering, cap))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.ray_index, cap*8))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.group, cap*4))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.bounce, cap*4))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.face_id, cap*4))) return rc;
    if ((rc = eng->alloc((void**)&r->hb.medium, cap*4))) return rc;
  }
  if ((rc = eng->alloc((void**)&r->dcounters, sizeof(Counters)))) return rc;
  CU(cudaMemsetAsync(r->dcounters, 0, sizeof(Counters), eng->stream));
  if (cfg->n_binnings > 0) {
    if (!cfg->binnings) return fail(ODW_EINVAL, "n_binnings > 0 but binnings is NULL");
    size_t off = 0;
    for (int b = 0; b < cfg->n_binnings; ++b) {
      const odw_binning& s = cfg->binnings[b];
      if (s.nu <= 0 || s.nv <= 0 || !(s.u_hi > s.u_lo) || !(s.v_hi > s.v_lo) || s.group < 0 || s.group >= sc->n_groups)
        return fail(ODW_EINVAL, "binning " + std::to_string(b) + ": bad specification");
      DBinning d{};
      for (int k = 0; k < 3; ++k) { d.origin[k] = s.origin[k]; d.ua[k] = s.uaxis[k]; d.va[k] = s.vaxis[k]; }
      d.u_lo = s.u_lo; d.v_lo = s.v_lo; d.u_hi = s.u_hi; d.v_hi = s.v_hi;
      d.u_scale = s.nu/(s.u_hi - s.u_lo); d.v_scale = s.nv/(s.v_hi - s.v_lo);
      d.group = s.group; d.nu = s.nu; d.nv = s.nv; d.weighted = s.weighted; d.offset = off;
      off += (size_t)s.nu*(size_t)s.nv;
      r->binnings.push_back(d);
    }
    r->total_bins = off;
    if ((rc = eng->alloc((void**)&r->dbins, off*sizeof(double)))) return rc;
    if ((rc = eng->alloc((void**)&r->dbinnings, r->binnings.size()*sizeof(DBinning)))) return rc;
    CU(cudaMemsetAsync(r->dbins, 0, off*sizeof(double), eng->stream));
    CU(cudaMemcpyAsync(r->dbinnings, r->binnings.data(), r->binnings.size()*sizeof(DBinning), cudaMemcpyHostToDevice, eng->stream));
  }
  memset(&p, 0, sizeof p);
  p.scene = sc->d;
  p.hits = r->hb;
  p.counters = r->dcounters;
  p.binnings = r->dbinnings; p.bins = r->dbins; p.n_binnings = cfg->n_binnings;
  p.n_rays = n_rays;
  p.max_len = cfg->max_ray_length; p.tol = std::max(cfg->dist_tol, 1e-6); p.power_tol = cfg->power_tol;
  p.max_isect = cfg->max_intersections; p.sequential = cfg->sequential;
  p.record_all = cfg->record_all_hits; p.store_hits = cfg->store_hits;
  return ODW_OK;
}

// Widening of the fp32 culling boxes (shells, BVH nodes).  The slab test runs in fp32 on the ray origin rounded to
// fp32: origin rounding <= 2^-24 |s|, each slab distance carries <= 3 roundings (2^-22 relative) of |b - s| <= E and
// of t <= max_len.  A box widened by distTol + 2e-6 E (about 8x that bound) therefore contains every point the exact
// fp64 test of the reference rule (box enlarged by distTol, ray.py:353-364) would accept.
static void set_cull_margin(TraceParams& p, const odw_scene* sc, double origin_bound) {
  const double E = sc->extent + origin_bound + p.max_len;
  p.cull_margin = std::nextafter((float)(p.tol + 2e-6*E), INFINITY);
  p.origin_bound = (float)origin_bound;
}

// The BVH boxes on the device carry the culling margin of the launch; re-widen them when it changes (tolerance,
// ray length or source changed).  Stream-ordered: earlier launches on the engine stream finish before the copy runs.
static int ensure_bvh_margin(odw_scene* sc, float margin) {
  if (!sc->use_bvh || sc->bvh_margin == margin) return ODW_OK;
  sc->bvh_staging = sc->bvh_host;
  for (BvhNode2& w : sc->bvh_staging)
    for (int a = 0; a < 3; ++a) {
      if (w.count[0] >= 0) { w.lo0[a] = std::nextafter(w.lo0[a] - margin, -INFINITY); w.hi0[a] = std::nextafter(w.hi0[a] + margin, INFINITY); }
      if (w.count[1] >= 0) { w.lo1[a] = std::nextafter(w.lo1[a] - margin, -INFINITY); w.hi1[a] = std::nextafter(w.hi1[a] + margin, INFINITY); }
    }
  CU(cudaMemcpyAsync(sc->bvh_dev, sc->bvh_staging.data(), sc->bvh_staging.size()*sizeof(BvhNode2), cudaMemcpyHostToDevice, sc->eng->stream));
  if (sc->bvh4_dev) {
    sc->bvh4_staging = sc->bvh4_host;
    for (Bvh4Node& w : sc->bvh4_staging)
      for (int k = 0; k < 4; ++k) {
        if (w.b[0][k] > w.b[1][k]) continue;                       // empty slot
        for (int a = 0; a < 3; ++a) { w.b[2*a][k] = std::nextafter(w.b[2*a][k] - margin, -INFINITY); w.b[2*a+1][k] = std::nextafter(w.b[2*a+1][k] + margin, INFINITY); }
      }
    CU(cudaMemcpyAsync(sc->bvh4_dev, sc->bvh4_staging.data(), sc->bvh4_staging.size()*sizeof(Bvh4Node), cudaMemcpyHostToDevice, sc->eng->stream));
  }
  CU(cudaStreamSynchronize(sc->eng->stream));      // the staging vectors are pageable and reused
  sc->bvh_margin = margin;
  return ODW_OK;
}

// IgnoredOpticalElements as a 256-bit mask; a group index the mask cannot hold is an error, never silently not ignored
static int set_ignore(TraceParams& p, const int32_t* ign, int n) {
  for (int i = 0; i < n; ++i) {
    if (ign[i] < 0) continue;
    if (ign[i] >= 256) return fail(ODW_EUNSUPPORTED, "ignored optical group index " + std::to_string(ign[i]) + " >= 256");
    p.ignore_mask[ign[i] >> 6] |= 1ull << (ign[i] & 63);
  }
  return ODW_OK;
}

// One wave of the wavefront formulation (BVH scenes, see odw_wavefront.cu): generate, then per bounce traverse + interact
// with the survivor count read back after every bounce (it sizes the next launches and ends the loop); once fewer than
// `tail` rays are left they finish in one launch.
static int run_wavefront_wave(odw_engine* eng, const TraceParams& q, bool mc, float bound, int need, uint64_t* launches) {
  const unsigned int n0 = (unsigned int)q.n_rays;
  if (n0 == 0) return ODW_OK;
  const size_t cap = n0;
  void *pool_a = nullptr, *pool_b = nullptr, *pool_c = nullptr, *hits = nullptr, *sort_temp = nullptr; unsigned int* ctr = nullptr;
  unsigned int *keys_out = nullptr, *iota = nullptr, *order = nullptr;
  int rc;
  auto cleanup = [&]() { eng->release(pool_a); eng->release(pool_b); eng->release(pool_c); eng->release(hits); eng->release(ctr);
                         eng->release(sort_temp); eng->release(keys_out); eng->release(iota); eng->release(order); };
  if ((rc = eng->alloc(&pool_a, cap*odw_wf_pool_bytes_per_ray())) || (rc = eng->alloc(&pool_b, cap*odw_wf_pool_bytes_per_ray())) ||
      (rc = eng->alloc(&hits, cap*16)) || (rc = eng->alloc((void**)&ctr, 16))) { cleanup(); return rc; }
  // Coherence sort: before the first traversal the rays are ordered by (origin cell, direction) so that the lanes of a warp
  // walk the same nodes (hugeArray, 2^24 rays: first traversal 7.5 -> 3.4 ms at 0.6 ms for the sort, 12 -> 24 of 32 lanes
  // active).  The interaction runs in the same order, so the survivors land in the next pool roughly ordered and the later
  // bounces inherit most of the coherence (second traversal 2.5 -> 1.5 ms); sorting those again gained nothing
  // (ODW_WF_SORT_BOUNCES, default 1; ODW_WF_SORT=0 switches the sort off; below ODW_WF_SORT_MIN rays it is skipped).
  unsigned int sort_min = 1u << 15; int sort_bounces = 1; size_t temp_bytes = 0;
  if (const char* w = getenv("ODW_WF_SORT")) { if (atoi(w) == 0) sort_bounces = 0; }
  if (const char* w = getenv("ODW_WF_SORT_BOUNCES")) { if (sort_bounces > 0) sort_bounces = std::max(0, atoi(w)); }
  if (const char* w = getenv("ODW_WF_SORT_MIN")) { long long v = atoll(w); if (v > 0) sort_min = (unsigned int)v; }
  if (sort_bounces > 0 && n0 >= sort_min) {
    cudaError_t es = odw_wf_sort(nullptr, &temp_bytes, pool_a, cap, nullptr, nullptr, nullptr, n0, eng->stream);
    if (es != cudaSuccess) { cleanup(); return fail(ODW_ECUDA, std::string("wavefront sort: ") + cudaGetErrorString(es)); }
    if ((rc = eng->alloc(&sort_temp, std::max<size_t>(temp_bytes, 16))) || (rc = eng->alloc((void**)&keys_out, cap*4)) ||
        (rc = eng->alloc((void**)&iota, cap*4)) || (rc = eng->alloc((void**)&order, cap*4)) ||
        (rc = eng->alloc(&pool_c, cap*odw_wf_pool_bytes_per_ray()))) { cleanup(); return rc; }
    if ((es = odw_wf_iota(iota, n0, eng->stream)) != cudaSuccess) { cleanup(); return fail(ODW_ECUDA, std::string("wavefront sort: ") + cudaGetErrorString(es)); }
  } else sort_bounces = 0;
  unsigned int tail = 65536;             // measured on hugeArray: 2048 / 8192 / 65536 -> 4.55 / 4.83 / 4.87e9 segments/s
  if (const char* w = getenv("ODW_WAVEFRONT_TAIL")) { long long v = atoll(w); if (v >= 0) tail = (unsigned int)v; }
  const int blocks = eng->sm_count*std::max(1, odw_wf_traverse_occupancy(q.scene.n_bvh_nodes));
  cudaStream_t st = eng->stream;
  unsigned int* host_n = reinterpret_cast<unsigned int*>(&eng->pinned_counters[2]);     // page-locked scratch of its own: [0], [1] hold chunk counters the host may not have read yet
  const bool wide = q.scene.bvh4 != nullptr;                                            // 4-wide traversal (wf_traverse4)
  // Monte-Carlo rays of a sorted first bounce are drawn twice: once for their coherence key, once by the traversal in coherence
  // order — cheaper than storing them and gathering five 16-byte columns per ray from unrelated addresses
  bool regen = wide && mc && sort_bounces > 0 && q.max_isect > 0 && n0 > tail;
  if (const char* w = getenv("ODW_WF_REGEN")) regen = regen && atoi(w) != 0;
  cudaError_t e = regen ? odw_wf_generate_keys(&q, pool_a, cap, bound, n0, st) : odw_wf_generate(&q, mc, pool_a, cap, bound, n0, st);
  if (launches) ++*launches;
  unsigned int n = q.max_isect > 0 ? n0 : 0;
  void *cur = pool_a, *nxt = pool_b;
  for (int bounce = 0; e == cudaSuccess && n > 0; ++bounce) {
    if (n <= tail) { e = odw_wf_tail(&q, mc, cur, cap, n, bounce, st); if (launches) ++*launches; break; }
    if ((e = cudaMemsetAsync(ctr, 0, 16, st)) != cudaSuccess) break;                    // ctr[0] = survivors, ctr[1] = fetch counter
    const bool sorted = bounce < sort_bounces && n >= sort_min;
    if (sorted) { if ((e = odw_wf_sort(sort_temp, &temp_bytes, cur, cap, keys_out, iota, order, n, st)) != cudaSuccess) break; }
    // sorted: the traversal moves every ray to its place in the ordered pool (pool_c) and the interaction reads that one
    if (wide) {
      if ((e = odw_wf_traverse4(&q, cur, cap, hits, n, ctr + 1, sorted ? order : nullptr, pool_c, need, (regen && bounce == 0 && sorted) ? 1 : 0,
                                eng->sm_count, st)) != cudaSuccess) break;
    } else
    if ((e = odw_wf_traverse(&q, cur, cap, hits, n, ctr + 1, sorted ? order : nullptr, pool_c, need, blocks, st)) != cudaSuccess) break;
    if ((e = odw_wf_interact(&q, mc, sorted ? pool_c : cur, hits, nxt, cap, bound, n, ctr, bounce, bounce + 1 < sort_bounces ? 1 : 0, st)) != cudaSuccess) break;
    if (launches) *launches += 2;
    if ((e = cudaMemcpyAsync(host_n, ctr, sizeof(unsigned int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    n = *host_n;
    std::swap(cur, nxt);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);                                  // the pools go back to the allocator
  cleanup();
  if (e != cudaSuccess) return fail(ODW_ECUDA, std::string("wavefront trace: ") + cudaGetErrorString(e));
  return ODW_OK;
}

// Issues the trace of p.n_rays rays as back-to-back launches on the engine stream (no synchronisation).
// features the launch needs (FEAT_* of odw_trace.cuh): the kernel instance is chosen from these
static int launch_features(const odw_scene* sc, const TraceParams& p, bool mc) {
  return (sc->ext_optics ? 1 : 0) | ((mc && p.src.kind == ODW_SRC_SURFACE) ? 2 : 0) | (p.sequential ? 4 : 0) | (p.n_binnings > 0 ? 8 : 0);
}

static int launch_waves(odw_engine* eng, const odw_scene* sc, const TraceParams& p, bool mc, uint64_t* launches) {
  const int need = launch_features(sc, p, mc);
  const bool presample = mc && p.src.kind == ODW_SRC_SURFACE && !sc->use_bvh && getenv("ODW_NO_PRESAMPLE") == nullptr;   // see below
  int per_sm = odw_trace_occupancy(mc && !presample, sc->use_bvh, need, sc->smem);
  if (per_sm <= 0) { cudaError_t e = cudaGetLastError(); return fail(ODW_ECUDA, std::string("trace kernel cannot be resident: ") + cudaGetErrorString(e)); }
  // persistent grid: a multiple of the SM count, no more blocks than there is work
  const int blocks = eng->sm_count*per_sm;
  // Waves: a long request is issued as launches of `wave` rays each, rotating over 4 streams so that the drain of one
  // wave (its last rays keep a few warps busy) overlaps the head of the next.  Inside a wave the warps claim rays
  // dynamically (TraceParams::ray_counter), so a wave has no fixed-assignment imbalance and its size is uncritical from
  // 2^19 up (wave size x stream count sweep in profiles/README.md: 29.5-29.9 ms per 1e8 rays for 2^19..2^23 rays on 2-8
  // streams, 30.5 ms for one stream of 2^23-ray waves; with the earlier fixed lane -> ray stride the optimum was 2^18 rays on 4
  // streams and one stream cost 50 % more).
  uint64_t wave = sc->use_bvh ? (1ull << 25)                   // BVH scenes: big waves amortise the per-bounce host round trip and the
                              : (1ull << 21);                  // thin last bounces (hugeArray: 4.59e9 at 2^24, 4.83e9 at 2^25 rays per wave)
  if (const char* w = getenv("ODW_RAYS_PER_LAUNCH")) { long long v = atoll(w); if (v > 0) wave = (uint64_t)v; }
  wave = std::min<uint64_t>(wave, 1ull << 31);
  wave = std::max<uint64_t>(wave, (p.n_rays + odw_engine::MAX_WAVES - 1)/odw_engine::MAX_WAVES);   // one claim counter per wave
  // Hit append and counters are atomic, so overlapping waves of one request are safe.
  int n_streams = odw_engine::DEFAULT_WAVE_STREAMS;
  if (const char* w = getenv("ODW_STREAMS")) n_streams = std::max(1, std::min(odw_engine::MAX_WAVE_STREAMS, atoi(w)));
  if ((sc->use_bvh && sc->wavefront) || p.n_rays <= wave) n_streams = 1;
  // The claim counters are zeroed on the engine stream BEFORE the fork event is recorded, so every wave stream is
  // ordered after the memset (and, through the join of the previous request, after every kernel that used the counters).
  if (!(sc->use_bvh && sc->wavefront) && p.n_rays > 0)
    CU(cudaMemsetAsync(eng->wave_counters, 0, ((p.n_rays + wave - 1)/wave)*sizeof(unsigned long long), eng->stream));
  // Surface sources: the emission code (face pick, area-uniform point, trim test, direction) is as large as the bounce loop, and a
  // warp of the register-resident kernel runs it whenever one of its rays ends — with rays of two or three segments the
  // kernel's instruction working set no longer fits the instruction cache (profiles/r02_v3_lines_lambertSource.txt: 58 % of the
  // stall samples are "no instruction").  The rays of a wave are therefore drawn by the sampling kernel first (all lanes in the
  // same code) and traced as an explicit list; 48 B per ray through HBM is nothing against that.
  // The pre-sampled rays of a wave are traced in coherence order (ray_sort_key: origin cell, direction cell): a surface emits
  // into a hemisphere from every point, so rays with neighbouring numbers have nothing in common and the lanes of a warp
  // cull different shells and accept different faces (lambert-source: 10 of 32 lanes in the face tests).
  // ODW_PRESAMPLE_SORT="begin,end": key bits sorted on, "0" = trace in ray order.
  int sort_b = ODW_PRESAMPLE_SORT_BEGIN, sort_e = ODW_PRESAMPLE_SORT_END;
  if (const char* w = getenv("ODW_PRESAMPLE_SORT")) { sort_b = sort_e = 0; sscanf(w, "%d,%d", &sort_b, &sort_e); }
  const bool sort_rays = presample && sort_e > sort_b;
  const uint64_t wave_cap = std::min<uint64_t>(wave, p.n_rays);
  double* sample_buf[odw_engine::MAX_WAVE_STREAMS] = {};
  unsigned int* sort_buf[odw_engine::MAX_WAVE_STREAMS] = {};     // keys, sorted keys, order: wave_cap entries each
  void* sort_temp[odw_engine::MAX_WAVE_STREAMS] = {};
  unsigned int* iota = nullptr;
  size_t temp_bytes = 0;
  auto release_buffers = [&]() {
    for (int k = 0; k < odw_engine::MAX_WAVE_STREAMS; ++k) { eng->release(sample_buf[k]); eng->release(sort_buf[k]); eng->release(sort_temp[k]); }
    eng->release(iota);
  };
  if (presample) {
    if (sort_rays) {
      cudaError_t es = odw_sort_pairs(nullptr, &temp_bytes, nullptr, nullptr, nullptr, nullptr, (unsigned int)wave_cap, sort_b, sort_e, eng->stream);
      if (es != cudaSuccess) return fail(ODW_ECUDA, std::string("pre-sample sort: ") + cudaGetErrorString(es));
      int rc = eng->alloc((void**)&iota, (size_t)wave_cap*4);
      if (rc) return rc;
      if ((es = odw_wf_iota(iota, (unsigned int)wave_cap, eng->stream)) != cudaSuccess) { release_buffers(); return fail(ODW_ECUDA, std::string("pre-sample sort: ") + cudaGetErrorString(es)); }
    }
    for (int i = 0; i < n_streams; ++i) {
      int rc = eng->alloc((void**)&sample_buf[i], (size_t)wave_cap*48);
      if (!rc && sort_rays) rc = eng->alloc((void**)&sort_buf[i], (size_t)wave_cap*12);
      if (!rc && sort_rays) rc = eng->alloc(&sort_temp[i], std::max<size_t>(temp_bytes, 16));
      if (rc) { release_buffers(); return rc; }
    }
  }
  // fork: after the memset of the claim counters and the iota fill, both on the engine stream
  if (n_streams > 1) {
    { cudaError_t e = cudaEventRecord(eng->ev_fork, eng->stream); if (e != cudaSuccess) { release_buffers(); return fail(ODW_ECUDA, cudaGetErrorString(e)); } }
    for (int i = 1; i < n_streams; ++i) CU(cudaStreamWaitEvent(eng->wave_stream[i], eng->ev_fork, 0));
  }
  uint64_t wave_index = 0;
  for (uint64_t off = 0; off < p.n_rays; off += wave, ++wave_index) {
    TraceParams q = p;
    q.ray_counter = eng->wave_counters + wave_index;
    q.n_rays = std::min<uint64_t>(wave, p.n_rays - off);
    q.first_ray = p.first_ray + off;
    if (!mc) {
      q.in_origins = p.in_origins + 3*off; q.in_dirs = p.in_dirs + 3*off;
      if (p.in_powers) q.in_powers = p.in_powers + off;
      if (p.out_nseg) q.out_nseg = p.out_nseg + off;
      if (p.out_final_point) q.out_final_point = p.out_final_point + 3*off;
      if (p.out_final_power) q.out_final_power = p.out_final_power + off;
      if (p.out_final_medium) q.out_final_medium = p.out_final_medium + off;
    }
    if (sc->use_bvh && sc->wavefront) {
      int rc = run_wavefront_wave(eng, q, mc, (float)std::max(1e-3, std::max(sc->extent, (double)q.origin_bound)), need, launches);
      if (rc) return rc;
      continue;
    }
    const uint64_t tpb = (uint64_t)odw_trace_threads();
    uint64_t want_w = (q.n_rays + tpb - 1)/tpb;
    int blocks_w = (int)std::min<uint64_t>((uint64_t)blocks, std::max<uint64_t>(1, want_w));
    cudaStream_t wst = eng->wave_stream[wave_index % (uint64_t)n_streams];
    if (presample) {
      double* buf = sample_buf[wave_index % (uint64_t)n_streams];            // waves of one stream run one after the other: the buffer is free again
      const int sblocks = (int)std::min<uint64_t>((uint64_t)eng->sm_count*8, std::max<uint64_t>(1, (q.n_rays + 255)/256));
      unsigned int* sb = sort_buf[wave_index % (uint64_t)n_streams];
      CU(odw_launch_sample(&q.src, q.seed, q.first_ray, q.n_rays, nullptr, nullptr, buf, buf + 3*q.n_rays, sort_rays ? sb : nullptr,
                           (float)std::max(1e-3, std::max(sc->extent, (double)q.origin_bound)), sblocks, wst));
      q.in_origins = buf; q.in_dirs = buf + 3*q.n_rays;
      if (sort_rays) {
        size_t tb = temp_bytes;
        CU(odw_sort_pairs(sort_temp[wave_index % (uint64_t)n_streams], &tb, sb, sb + wave_cap, iota, sb + 2*wave_cap, (unsigned int)q.n_rays, sort_b, sort_e, wst));
        q.in_order = sb + 2*wave_cap;
      }
      CU(odw_launch_trace(&q, false, false, need, blocks_w, sc->smem, wst));
      if (launches) *launches += 2;
      continue;
    }
    CU(odw_launch_trace(&q, mc, sc->use_bvh, need, blocks_w, sc->smem, wst));
    if (launches) ++*launches;
  }
  release_buffers();   // stream order protects them: every later user is queued after the join below
  for (int i = 1; i < n_streams; ++i) { CU(cudaEventRecord(eng->ev_join[i], eng->wave_stream[i])); CU(cudaStreamWaitEvent(eng->stream, eng->ev_join[i], 0)); }
  return ODW_OK;
}

static int run_trace(odw_engine* eng, const odw_scene* sc, odw_result* r, const TraceParams& p, bool mc) {
  uint64_t launches = 0;
  CU(cudaEventRecord(eng->ev0, eng->stream));
  int rc = launch_waves(eng, sc, p, mc, &launches);
  if (rc) return rc;
  CU(cudaEventRecord(eng->ev1, eng->stream));
  Counters c;
  CU(cudaMemcpyAsync(&c, r->dcounters, sizeof c, cudaMemcpyDeviceToHost, eng->stream));
  CU(cudaStreamSynchronize(eng->stream));
  CU(cudaEventElapsedTime(&r->ms, eng->ev0, eng->ev1));
  r->counts.rays = p.n_rays; r->counts.segments = c.segments; r->counts.hits = c.hits;
  r->counts.hits_dropped = c.hits_dropped; r->counts.escaped = c.escaped; r->counts.depth_terminated = c.depth_terminated;
  r->counts.waves = launches;
  r->counts.sm_clock_khz = c.dbg_ns ? (uint64_t)((double)c.dbg_cycles*1e6/(double)c.dbg_ns) : 0;
  return c.hits_dropped ? fail(ODW_EOVERFLOW, std::to_string(c.hits_dropped) + " hits did not fit hit_capacity " + std::to_string(p.hits.capacity)) : ODW_OK;
}

extern "C" int odw_trace_mc(odw_scene* sc, odw_source* src, const odw_trace_cfg* cfg, uint64_t seed, uint64_t first_ray,
                            uint64_t n_rays, odw_result** out) {
  if (!sc || !src || !cfg || !out) return fail(ODW_EINVAL, "odw_trace_mc: NULL argument");
  *out = nullptr;
  if (sc->eng != src->eng) return fail(ODW_EINVAL, "odw_trace_mc: scene and source belong to different engines");
  odw_engine* eng = sc->eng;
  CU(cudaSetDevice(eng->device));
  TraceParams p;
  int rc = prepare_result(eng, sc, cfg, n_rays, out, p);
  if (rc) { odw_result_destroy(*out); *out = nullptr; return rc; }
  p.src = src->d;
  p.seed = seed; p.first_ray = first_ray;
  p.max_len = cfg->max_ray_length*src->max_ray_length_scale;                 // ray.py:48-53
  p.max_isect = (int)(cfg->max_intersections*src->max_intersections_scale);
  p.wavelength = src->d.wavelength;
  if ((rc = set_ignore(p, src->ignored.data(), (int)src->ignored.size()))) { odw_result_destroy(*out); *out = nullptr; return rc; }
  set_cull_margin(p, sc, src->origin_bound);
  if ((rc = ensure_bvh_margin(sc, p.cull_margin))) { odw_result_destroy(*out); *out = nullptr; return rc; }
  rc = run_trace(eng, sc, *out, p, true);
  if (rc && rc != ODW_EOVERFLOW) { odw_result_destroy(*out); *out = nullptr; }
  return rc;
}

// Monte-Carlo trace with HOST result buffers: the ray range is processed in chunks; while chunk c+1 is traced on the
// compute stream, the hit columns of chunk c go device->host on the copy stream (double-buffered device hit lists).
// This is the call the plugin's runSimulationIteration replacement makes when it wants hit lists on the host
// (reference results_store.py:641-648 appends to Python lists; here the rows land in the caller's arrays).
extern "C" int odw_trace_mc_host(odw_scene* sc, odw_source* src, const odw_trace_cfg* cfg, uint64_t seed, uint64_t first_ray,
                                 uint64_t n_rays, const odw_hits_view* host, uint64_t* n_hits_out, odw_counts* counts_out) {
  if (!sc || !src || !cfg || !host) return fail(ODW_EINVAL, "odw_trace_mc_host: NULL argument");
  if (sc->eng != src->eng) return fail(ODW_EINVAL, "odw_trace_mc_host: scene and source belong to different engines");
  if (cfg->n_binnings > 0) return fail(ODW_EUNSUPPORTED, "odw_trace_mc_host: use odw_trace_mc for device binning");
  odw_engine* eng = sc->eng;
  CU(cudaSetDevice(eng->device));
  uint64_t chunk = sc->use_bvh ? (1ull << 25) : (1ull << 23);   // BVH scenes: a chunk is one wave of the wavefront path, which wants to be large
  if (const char* w = getenv("ODW_HOST_CHUNK")) { long long v = atoll(w); if (v > 0) chunk = (uint64_t)v; }
  chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(n_rays, 1));
  // cfg->hit_capacity = rows the caller expects for the WHOLE range (0: two per ray); the per-chunk device lists get the
  // same rows-per-ray ratio, so a caller that repeats the call with a larger capacity after ODW_EOVERFLOW (scenes that
  // record more than two hits per ray: transparent detectors, record_all_hits) also gets larger device lists.  The chunk
  // shrinks when that would take more than 2^27 rows per list (81 B per row, two lists).
  uint64_t rows_per_ray = 2;
  if (cfg->hit_capacity && n_rays) rows_per_ray = std::max<uint64_t>(1, (cfg->hit_capacity + n_rays - 1)/n_rays);
  if (cfg->max_intersections > 0) rows_per_ray = std::min<uint64_t>(rows_per_ray, (uint64_t)cfg->max_intersections*std::max(1.0, src->max_intersections_scale) + 1);
  while (chunk > (1ull << 16) && chunk*rows_per_ray > (1ull << 27)) chunk >>= 1;
  odw_trace_cfg ccfg = *cfg;
  ccfg.store_hits = 1;
  ccfg.hit_capacity = std::max<uint64_t>(1024, rows_per_ray*chunk);
  odw_result* r[2] = {nullptr, nullptr};
  TraceParams p[2];
  auto cleanup = [&]() { odw_result_destroy(r[0]); odw_result_destroy(r[1]); };
  for (int b = 0; b < 2; ++b) {
    int rc = prepare_result(eng, sc, &ccfg, chunk, &r[b], p[b]);
    if (rc) { cleanup(); return rc; }
    p[b].src = src->d; p[b].seed = seed;
    p[b].max_len = cfg->max_ray_length*src->max_ray_length_scale;
    p[b].max_isect = (int)(cfg->max_intersections*src->max_intersections_scale);
    p[b].wavelength = src->d.wavelength;
    if ((rc = set_ignore(p[b], src->ignored.data(), (int)src->ignored.size()))) { cleanup(); return rc; }
    set_cull_margin(p[b], sc, src->origin_bound);
    if ((rc = ensure_bvh_margin(sc, p[b].cull_margin))) { cleanup(); return rc; }
  }
  const uint64_t n_chunks = (n_rays + chunk - 1)/chunk;
  odw_counts total; memset(&total, 0, sizeof total);
  total.rays = n_rays;
  uint64_t stored = 0, dropped_host = 0, launches = 0;
  auto issue = [&](uint64_t c) -> int {        // queue the trace of chunk c on the compute stream
    const int b = (int)(c & 1);
    if (c >= 2) CU(cudaStreamWaitEvent(eng->stream, eng->ev_copy[b], 0));   // buffer b still being copied out
    CU(cudaMemsetAsync(r[b]->dcounters, 0, sizeof(Counters), eng->stream));
    p[b].first_ray = first_ray + c*chunk;
    p[b].n_rays = std::min<uint64_t>(chunk, n_rays - c*chunk);
    int rc = launch_waves(eng, sc, p[b], true, &launches);
    if (rc) return rc;
    CU(cudaMemcpyAsync(&eng->pinned_counters[b], r[b]->dcounters, sizeof(Counters), cudaMemcpyDeviceToHost, eng->stream));
    CU(cudaEventRecord(eng->ev_trace[b], eng->stream));
    return ODW_OK;
  };
  int rc = ODW_OK;
  if (n_chunks > 0 && (rc = issue(0))) { cleanup(); return rc; }
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const int b = (int)(c & 1);
    if (c + 1 < n_chunks && (rc = issue(c + 1))) { cleanup(); return rc; }
    CU(cudaEventSynchronize(eng->ev_trace[b]));
    const Counters k = eng->pinned_counters[b];
    total.segments += k.segments; total.hits += k.hits; total.hits_dropped += k.hits_dropped;
    total.escaped += k.escaped; total.depth_terminated += k.depth_terminated;
    if (k.dbg_ns) total.sm_clock_khz = (uint64_t)((double)k.dbg_cycles*1e6/(double)k.dbg_ns);
    uint64_t have = std::min<uint64_t>(k.hits, r[b]->hb.capacity);
    uint64_t n = std::min<uint64_t>(have, host->capacity > stored ? host->capacity - stored : 0);
    dropped_host += have - n;
    cudaStream_t cs = eng->copy_stream;
    const HitBuffers& hb = r[b]->hb;
    if (n) {
      if (host->points)      CU(cudaMemcpyAsync(host->points + 3*stored, hb.points, n*24, cudaMemcpyDeviceToHost, cs));
      if (host->directions)  CU(cudaMemcpyAsync(host->directions + 3*stored, hb.dirs, n*24, cudaMemcpyDeviceToHost, cs));
      if (host->powers)      CU(cudaMemcpyAsync(host->powers + stored, hb.powers, n*8, cudaMemcpyDeviceToHost, cs));
      if (host->is_entering) CU(cudaMemcpyAsync(host->is_entering + stored, hb.entering, n, cudaMemcpyDeviceToHost, cs));
      if (host->ray_index)   CU(cudaMemcpyAsync(host->ray_index + stored, hb.ray_index, n*8, cudaMemcpyDeviceToHost, cs));
      if (host->group)       CU(cudaMemcpyAsync(host->group + stored, hb.group, n*4, cudaMemcpyDeviceToHost, cs));
      if (host->bounce)      CU(cudaMemcpyAsync(host->bounce + stored, hb.bounce, n*4, cudaMemcpyDeviceToHost, cs));
      if (host->face_id)     CU(cudaMemcpyAsync(host->face_id + stored, hb.face_id, n*4, cudaMemcpyDeviceToHost, cs));
      if (host->medium)      CU(cudaMemcpyAsync(host->medium + stored, hb.medium, n*4, cudaMemcpyDeviceToHost, cs));
    }
    CU(cudaEventRecord(eng->ev_copy[b], cs));
    stored += n;
  }
  CU(cudaStreamSynchronize(eng->copy_stream));
  CU(cudaStreamSynchronize(eng->stream));
  cleanup();
  total.waves = launches;
  total.hits_dropped += dropped_host;
  if (n_hits_out) *n_hits_out = stored;
  if (counts_out) *counts_out = total;
  return total.hits_dropped ? fail(ODW_EOVERFLOW, std::to_string(total.hits_dropped) + " hits did not fit the host / chunk hit buffers") : ODW_OK;
}

extern "C" int odw_sample_mc(odw_source* src, uint64_t seed, uint64_t first_ray, uint64_t n, double* first_var, double* phi,
                             double* origins, double* directions) {
  if (!src) return fail(ODW_EINVAL, "odw_sample_mc: NULL source");
  odw_engine* eng = src->eng;
  CU(cudaSetDevice(eng->device));
  double *df = nullptr, *dp = nullptr, *dorg = nullptr, *dd = nullptr;
  int rc = ODW_OK;
  if (first_var && (rc = eng->alloc((void**)&df, n*8))) return rc;
  if (phi && (rc = eng->alloc((void**)&dp, n*8))) return rc;
  if (origins && (rc = eng->alloc((void**)&dorg, n*24))) return rc;
  if (directions && (rc = eng->alloc((void**)&dd, n*24))) return rc;
  int blocks = (int)std::min<uint64_t>((uint64_t)eng->sm_count*8, std::max<uint64_t>(1, (n + 255)/256));
  if (n) CU(odw_launch_sample(&src->d, seed, first_ray, n, df, dp, dorg, dd, nullptr, 1.0f, blocks, eng->stream));
  if (first_var) CU(cudaMemcpyAsync(first_var, df, n*8, cudaMemcpyDeviceToHost, eng->stream));
  if (phi) CU(cudaMemcpyAsync(phi, dp, n*8, cudaMemcpyDeviceToHost, eng->stream));
  if (origins) CU(cudaMemcpyAsync(origins, dorg, n*24, cudaMemcpyDeviceToHost, eng->stream));
  if (directions) CU(cudaMemcpyAsync(directions, dd, n*24, cudaMemcpyDeviceToHost, eng->stream));
  CU(cudaStreamSynchronize(eng->stream));
  eng->release(df); eng->release(dp); eng->release(dorg); eng->release(dd);
  return ODW_OK;
}

extern "C" int odw_trace_rays(odw_scene* sc, const odw_trace_cfg* cfg, const double* origins, const double* directions,
                              const double* powers, const int32_t* ignored_groups, int32_t n_ignored, uint64_t n_rays,
                              odw_result** out) {
  if (!sc || !cfg || !out || (n_rays && (!origins || !directions))) return fail(ODW_EINVAL, "odw_trace_rays: NULL argument");
  *out = nullptr;
  odw_engine* eng = sc->eng;
  CU(cudaSetDevice(eng->device));
  TraceParams p;
  int rc = prepare_result(eng, sc, cfg, n_rays, out, p);
  if (rc) { odw_result_destroy(*out); *out = nullptr; return rc; }
  odw_result* r = *out;
  auto bail = [&](int code) { odw_result_destroy(r); *out = nullptr; return code; };
  if ((rc = eng->alloc((void**)&r->d_in_o, n_rays*24))) return bail(rc);
  if ((rc = eng->alloc((void**)&r->d_in_d, n_rays*24))) return bail(rc);
  if (powers && (rc = eng->alloc((void**)&r->d_in_p, n_rays*8))) return bail(rc);
  if ((rc = eng->alloc((void**)&r->d_nseg, n_rays*4))) return bail(rc);
  if ((rc = eng->alloc((void**)&r->d_final_point, n_rays*24))) return bail(rc);
  if ((rc = eng->alloc((void**)&r->d_final_power, n_rays*8))) return bail(rc);
  if ((rc = eng->alloc((void**)&r->d_final_medium, n_rays*4))) return bail(rc);
  if (n_rays) {
    CU(cudaMemcpyAsync(r->d_in_o, origins, n_rays*24, cudaMemcpyHostToDevice, eng->stream));
    CU(cudaMemcpyAsync(r->d_in_d, directions, n_rays*24, cudaMemcpyHostToDevice, eng->stream));
    if (powers) CU(cudaMemcpyAsync(r->d_in_p, powers, n_rays*8, cudaMemcpyHostToDevice, eng->stream));
  }
  p.in_origins = r->d_in_o; p.in_dirs = r->d_in_d; p.in_powers = powers ? r->d_in_p : nullptr;
  p.out_nseg = r->d_nseg; p.out_final_point = r->d_final_point; p.out_final_power = r->d_final_power; p.out_final_medium = r->d_final_medium;
  p.first_ray = 0;
  p.seed = cfg->scatter_seed; p.src.source_id = 0;       // Philox stream of the stochastic-surface draws of an explicit list
  p.wavelength = cfg->wavelength > 0 ? cfg->wavelength : 500.0;
  if ((rc = set_ignore(p, ignored_groups, ignored_groups ? n_ignored : 0))) return bail(rc);
  double origin_bound = 0;
  for (uint64_t i = 0; i < 3*n_rays; ++i) origin_bound = std::max(origin_bound, std::fabs(origins[i]));
  if (!std::isfinite(origin_bound)) return bail(fail(ODW_EINVAL, "odw_trace_rays: non-finite ray origin"));
  set_cull_margin(p, sc, origin_bound);
  if ((rc = ensure_bvh_margin(sc, p.cull_margin))) return bail(rc);
  rc = run_trace(eng, sc, r, p, false);
  if (rc && rc != ODW_EOVERFLOW) return bail(rc);
  return rc;
}

// ------------------------------------------------------------------------------------------
extern "C" int odw_result_counts(const odw_result* r, odw_counts* out) {
  if (!r || !out) return fail(ODW_EINVAL, "odw_result_counts: NULL argument");
  *out = r->counts;
  return ODW_OK;
}

extern "C" int odw_result_kernel_ms(const odw_result* r, double* ms) {
  if (!r || !ms) return fail(ODW_EINVAL, "odw_result_kernel_ms: NULL argument");
  *ms = r->ms;
  return ODW_OK;
}

extern "C" int odw_result_hits(const odw_result* r, odw_hits_view* v, int sorted, uint64_t* n_out) {
  if (!r || !v) return fail(ODW_EINVAL, "odw_result_hits: NULL argument");
  odw_engine* eng = r->eng;
  CU(cudaSetDevice(eng->device));
  uint64_t stored = std::min<uint64_t>(r->counts.hits, r->hb.capacity);
  uint64_t n = std::min<uint64_t>(stored, v->capacity);
  if (n_out) *n_out = n;
  if (n == 0) return ODW_OK;
  cudaStream_t st = eng->stream;
  if (!sorted) {
    if (v->points)      CU(cudaMemcpyAsync(v->points, r->hb.points, n*24, cudaMemcpyDeviceToHost, st));
    if (v->directions)  CU(cudaMemcpyAsync(v->directions, r->hb.dirs, n*24, cudaMemcpyDeviceToHost, st));
    if (v->powers)      CU(cudaMemcpyAsync(v->powers, r->hb.powers, n*8, cudaMemcpyDeviceToHost, st));
    if (v->is_entering) CU(cudaMemcpyAsync(v->is_entering, r->hb.entering, n, cudaMemcpyDeviceToHost, st));
    if (v->ray_index)   CU(cudaMemcpyAsync(v->ray_index, r->hb.ray_index, n*8, cudaMemcpyDeviceToHost, st));
    if (v->group)       CU(cudaMemcpyAsync(v->group, r->hb.group, n*4, cudaMemcpyDeviceToHost, st));
    if (v->bounce)      CU(cudaMemcpyAsync(v->bounce, r->hb.bounce, n*4, cudaMemcpyDeviceToHost, st));
    if (v->face_id)     CU(cudaMemcpyAsync(v->face_id, r->hb.face_id, n*4, cudaMemcpyDeviceToHost, st));
    if (v->medium)      CU(cudaMemcpyAsync(v->medium, r->hb.medium, n*4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ODW_OK;
  }
  // sorted by (ray_index, bounce): the append order of the kernel is not deterministic
  std::vector<unsigned long long> ray(stored); std::vector<int32_t> bounce(stored);
  CU(cudaMemcpyAsync(ray.data(), r->hb.ray_index, stored*8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(bounce.data(), r->hb.bounce, stored*4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::vector<uint64_t> order(stored);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
    return ray[a] != ray[b] ? ray[a] < ray[b] : bounce[a] < bounce[b]; });
  auto gather = [&](void* dst, const void* dsrc, size_t elem) -> int {
    if (!dst) return ODW_OK;
    std::vector<unsigned char> tmp(stored*elem);
    CU(cudaMemcpyAsync(tmp.data(), dsrc, stored*elem, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    unsigned char* o = static_cast<unsigned char*>(dst);
    for (uint64_t i = 0; i < n; ++i) memcpy(o + i*elem, tmp.data() + order[i]*elem, elem);
    return ODW_OK;
  };
  int rc;
  if ((rc = gather(v->points, r->hb.points, 24))) return rc;
  if ((rc = gather(v->directions, r->hb.dirs, 24))) return rc;
  if ((rc = gather(v->powers, r->hb.powers, 8))) return rc;
  if ((rc = gather(v->is_entering, r->hb.entering, 1))) return rc;
  if ((rc = gather(v->ray_index, r->hb.ray_index, 8))) return rc;
  if ((rc = gather(v->group, r->hb.group, 4))) return rc;
  if ((rc = gather(v->bounce, r->hb.bounce, 4))) return rc;
  if ((rc = gather(v->face_id, r->hb.face_id, 4))) return rc;
  if ((rc = gather(v->medium, r->hb.medium, 4))) return rc;
  return ODW_OK;
}

extern "C" int odw_result_histogram(const odw_result* r, int32_t b, double* bins_out) {
  if (!r || !bins_out) return fail(ODW_EINVAL, "odw_result_histogram: NULL argument");
  if (b < 0 || b >= (int)r->binnings.size()) return fail(ODW_EINVAL, "odw_result_histogram: binning index out of range");
  CU(cudaSetDevice(r->eng->device));
  const DBinning& d = r->binnings[(size_t)b];
  CU(cudaMemcpy(bins_out, r->dbins + d.offset, (size_t)d.nu*d.nv*sizeof(double), cudaMemcpyDeviceToHost));
  return ODW_OK;
}

extern "C" int odw_result_histogram_device(const odw_result* r, int32_t b, void** dptr, uint64_t* n_bins) {
  if (!r || !dptr) return fail(ODW_EINVAL, "odw_result_histogram_device: NULL argument");
  if (b < 0 || b >= (int)r->binnings.size()) return fail(ODW_EINVAL, "odw_result_histogram_device: binning index out of range");
  const DBinning& d = r->binnings[(size_t)b];
  *dptr = r->dbins + d.offset;
  if (n_bins) *n_bins = (uint64_t)d.nu*d.nv;
  return ODW_OK;
}

extern "C" int odw_result_ray_media(const odw_result* r, int32_t* final_medium) {
  if (!r || !final_medium) return fail(ODW_EINVAL, "odw_result_ray_media: NULL argument");
  if (!r->d_final_medium) return fail(ODW_EINVAL, "odw_result_ray_media: only available for odw_trace_rays results");
  CU(cudaSetDevice(r->eng->device));
  if (r->n_rays) CU(cudaMemcpy(final_medium, r->d_final_medium, r->n_rays*4, cudaMemcpyDeviceToHost));
  return ODW_OK;
}

extern "C" int odw_result_ray_summary(const odw_result* r, int32_t* n_segments, double* final_points, double* final_powers) {
  if (!r) return fail(ODW_EINVAL, "odw_result_ray_summary: NULL argument");
  if (!r->d_nseg) return fail(ODW_EINVAL, "odw_result_ray_summary: only available for odw_trace_rays results");
  CU(cudaSetDevice(r->eng->device));
  uint64_t n = r->n_rays;
  if (n == 0) return ODW_OK;
  if (n_segments) CU(cudaMemcpy(n_segments, r->d_nseg, n*4, cudaMemcpyDeviceToHost));
  if (final_points) CU(cudaMemcpy(final_points, r->d_final_point, n*24, cudaMemcpyDeviceToHost));
  if (final_powers) CU(cudaMemcpy(final_powers, r->d_final_power, n*8, cudaMemcpyDeviceToHost));
  return ODW_OK;
}
