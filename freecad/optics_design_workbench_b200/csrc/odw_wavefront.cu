// odw_wavefront.cu — wavefront trace kernels for scenes that need the BVH (hundreds of faces and more).
//
// Why a second formulation: in the register-resident kernel (odw_kernels.cu) a warp does one bounce per iteration and
// waits for its slowest lane.  With a BVH the traversal lengths of the 32 rays of a warp differ wildly (a ray that misses
// everything visits a handful of nodes, one threading a lattice of spheres hundreds): on benchmark/hugeArray.FCStd only
// 4 of 32 lanes were active on average (profiles/r01_v7_trace_hugeArray_raw.json).  Here the bounce loop of
// Ray.traceRay (reference freecad_elements/ray.py:91-281) is cut at findNearestIntersection:
//
//   wf_generate   rays of the wave -> ray pool in HBM (SoA of 16-byte columns: coalesced 128-bit loads / stores)
//   per bounce:
//     wf_traverse   persistent warps; a lane whose traversal ended stores (t, face) and immediately takes the next ray
//                   of the pool (warp-aggregated atomic on the fetch counter), so every lane always traverses
//     wf_interact   one lane per ray: hit -> normal / Snell / mirror / grating / absorber, hit append, survivors are
//                   compacted into the other pool (warp ballot + one atomic per warp)
//   wf_tail       the last few thousand survivors of a wave finish in one launch (one lane per ray, whole bounce loop)
//
// HBM traffic per segment is the wavefront figure of SURVEY.md §8d (ray state read + written once per bounce).
#include <algorithm>
#include "odw_trace.cuh"
#include <cub/device/device_radix_sort.cuh>

// ---- ray pool ------------------------------------------------------------------------------
struct WfPool {
  double2* a0;       // (ox, oy)
  double2* a1;       // (oz, dx)
  double2* a2;       // (dy, dz)          d = unit direction
  double2* a3;       // (power, dscale)
  ulonglong2* a4;    // (ray number inside the launch, medium | seq_index << 32)
  unsigned int* key; // coherence key of the ray (origin cell, direction), see ray_sort_key
  float bound;       // origins are binned on a 16^3 grid over [-bound, bound]^3
};

// (coherence key: ray_sort_key, odw_device.cuh)
struct WfRay {
  double point[3], dn[3], power, dscale;
  unsigned long long i;
  int medium, seq_index;
};

template <bool WITH_KEY = true>
__device__ __forceinline__ void pool_store(const WfPool& pl, unsigned int slot, const double* point, const double* dn,
                                           double power, double dscale, unsigned long long i, int medium, int seq_index) {
  pl.a0[slot] = make_double2(point[0], point[1]);
  pl.a1[slot] = make_double2(point[2], dn[0]);
  pl.a2[slot] = make_double2(dn[1], dn[2]);
  pl.a3[slot] = make_double2(power, dscale);
  pl.a4[slot] = make_ulonglong2(i, (unsigned long long)(unsigned int)medium | ((unsigned long long)(unsigned int)seq_index << 32));
  if (WITH_KEY) pl.key[slot] = ray_sort_key(point, dn, pl.bound);
}

// ---- resumable BVH traversal, "while-while" form --------------------------------------------------------------
// Same rule as find_nearest_bvh (odw_trace.cuh), split so that the lanes of a warp spend their time in the same code:
// `cur` is the lane's next work item: an inner node (>= 0), a leaf (<= -2, first primitive and count packed) or
// TRAV_DONE.  The kernel first lets every lane descend through inner nodes (cheap fp32 box tests) until each holds a
// leaf or is done, then all lanes holding a leaf run the exact fp64 face tests together.  Testing a leaf the moment it
// is found (as find_nearest_bvh does) left 7 of 32 lanes active: every lane sat in a different part of the step.
#define TRAV_DONE (-1)
struct BvhTrav {       // scalars only: the two stack arrays are separate locals of the kernel, so that the dynamic indexing
  NearestHit h;        // they need does not drag cur / sp / limf into local memory with them
  float sx, sy, sz, ix, iy, iz, limf;
  int cur, sp;
};
struct BvhStack { int* ref; float* t; };

__device__ __forceinline__ int leaf_ref(int first, int count) { return -2 - ((first << 4) | count); }

__device__ __forceinline__ void bvh_begin(BvhTrav& tr, const TraceParams& p, const double* s, const double* dn) {
  const double tmax = p.max_len + p.tol;
  tr.h.tA = 1e300; tr.h.tB = 1e300; tr.h.lim = tmax; tr.h.fA = -1; tr.h.fB = -1;
  tr.sx = (float)s[0]; tr.sy = (float)s[1]; tr.sz = (float)s[2];
  tr.ix = __frcp_rn((float)dn[0]); tr.iy = __frcp_rn((float)dn[1]); tr.iz = __frcp_rn((float)dn[2]);
  tr.limf = (float)tmax*1.000002f;
  tr.cur = 0; tr.sp = 0;
}

// next stacked item that can still matter
__device__ __forceinline__ int bvh_pop(BvhTrav& tr, const BvhStack& st) {
  while (tr.sp > 0) {
    --tr.sp;
    if (st.t[tr.sp] <= tr.limf) return st.ref[tr.sp];
  }
  return TRAV_DONE;
}

// cur is an inner node: test both children, go to the nearer one, stack the other
// Nodes staged in shared memory: node n = 16-byte words 4n .. 4n+3.  Lanes of a warp read the SAME word of DIFFERENT nodes, and a
// 64-byte record offers only two bank positions per word, so the words are swizzled: word w lives in row w / 8 at position
// (w % 8) ^ (row % 8), which spreads the lanes over all eight 16-byte bank groups.
__device__ __forceinline__ int staged_word(int node, int k) {
  const int row = node >> 1, q = ((node & 1) << 2) | k;
  return (row << 3) | (q ^ (row & 7));
}

__device__ __forceinline__ void bvh_inner_step(BvhTrav& tr, const BvhStack& st, const TraceParams& p, const float4* s_nodes, int n_staged) {
  const float sx = tr.sx, sy = tr.sy, sz = tr.sz, ix = tr.ix, iy = tr.iy, iz = tr.iz;
  float4 a, b, c; int4 d;
  if (tr.cur < n_staged) {
    a = s_nodes[staged_word(tr.cur, 0)]; b = s_nodes[staged_word(tr.cur, 1)]; c = s_nodes[staged_word(tr.cur, 2)];
    const float4 dd = s_nodes[staged_word(tr.cur, 3)];
    d = make_int4(__float_as_int(dd.x), __float_as_int(dd.y), __float_as_int(dd.z), __float_as_int(dd.w));
  } else {
    const float4* q = reinterpret_cast<const float4*>(p.scene.bvh + tr.cur);
    a = __ldg(q); b = __ldg(q + 1); c = __ldg(q + 2);
    d = __ldg(reinterpret_cast<const int4*>(q + 3));
  }
  // child 0: lo = (a.x, a.y, a.z), hi = (a.w, b.x, b.y);  child 1: lo = (b.z, b.w, c.x), hi = (c.y, c.z, c.w)
  float ta = (a.x - sx)*ix, tb = (a.w - sx)*ix;
  float n0 = fminf(ta, tb), f0 = fmaxf(ta, tb);
  ta = (a.y - sy)*iy; tb = (b.x - sy)*iy;
  n0 = fmaxf(n0, fminf(ta, tb)); f0 = fminf(f0, fmaxf(ta, tb));
  ta = (a.z - sz)*iz; tb = (b.y - sz)*iz;
  n0 = fmaxf(n0, fminf(ta, tb)); f0 = fminf(f0, fmaxf(ta, tb));
  ta = (b.z - sx)*ix; tb = (c.y - sx)*ix;
  float n1 = fminf(ta, tb), f1 = fmaxf(ta, tb);
  ta = (b.w - sy)*iy; tb = (c.z - sy)*iy;
  n1 = fmaxf(n1, fminf(ta, tb)); f1 = fminf(f1, fmaxf(ta, tb));
  ta = (c.x - sz)*iz; tb = (c.w - sz)*iz;
  n1 = fmaxf(n1, fminf(ta, tb)); f1 = fminf(f1, fmaxf(ta, tb));
  const bool hit0 = d.z >= 0 && n0 <= f0 && f0 >= 0.0f && n0 <= tr.limf;
  const bool hit1 = d.w >= 0 && n1 <= f1 && f1 >= 0.0f && n1 <= tr.limf;
  const int r0 = d.z > 0 ? leaf_ref(d.x, d.z) : d.x, r1 = d.w > 0 ? leaf_ref(d.y, d.w) : d.y;
  if (hit0 && hit1) {
    const bool first0 = n0 <= n1;
    if (tr.sp < ODW_BVH_STACK) { st.ref[tr.sp] = first0 ? r1 : r0; st.t[tr.sp] = first0 ? n1 : n0; ++tr.sp; }
    tr.cur = first0 ? r0 : r1;
  } else if (hit0) tr.cur = r0;
  else if (hit1) tr.cur = r1;
  else tr.cur = bvh_pop(tr, st);
}

// cur is a leaf: exact fp64 tests of its faces
template <int FEAT>
__device__ __forceinline__ void bvh_leaf_step(BvhTrav& tr, const BvhStack& st, const TraceParams& p, const double* s, const double* dn, int medium, int seq_index) {
  const int packed = -2 - tr.cur, first = packed >> 4, count = packed & 15;
  const double tmax = p.max_len + p.tol;
  for (int k = 0; k < count; ++k) {
    const int fi = __ldg(p.scene.bvh_prims + first + k);
    test_face<true, FEAT>(p.scene.faces[fi], fi, p, s, dn, medium, seq_index, tmax, tr.h);
  }
  tr.limf = (float)tr.h.lim*1.000002f;
  tr.cur = bvh_pop(tr, st);
}

__device__ __forceinline__ void flush_counters(const TraceParams& p, const unsigned int* s_cnt) {
  if (s_cnt[CNT_SEGMENTS]) atomicAdd(&p.counters->segments, (unsigned long long)s_cnt[CNT_SEGMENTS]);
  if (!p.store_hits && s_cnt[CNT_HITS]) atomicAdd(&p.counters->hits, (unsigned long long)s_cnt[CNT_HITS]);
  if (s_cnt[CNT_DROPPED]) atomicAdd(&p.counters->hits_dropped, (unsigned long long)s_cnt[CNT_DROPPED]);
  if (s_cnt[CNT_ESCAPED]) atomicAdd(&p.counters->escaped, (unsigned long long)s_cnt[CNT_ESCAPED]);
  if (s_cnt[CNT_DEPTH]) atomicAdd(&p.counters->depth_terminated, (unsigned long long)s_cnt[CNT_DEPTH]);
}

// ---- kernels -------------------------------------------------------------------------------
// rays [0, n) of the launch -> pool slots [0, n)
template <bool MC>
__global__ void __launch_bounds__(256) wf_generate(const __grid_constant__ TraceParams p, WfPool pool, unsigned int n) {
  __shared__ unsigned int s_cnt[CNT_N];
  if (threadIdx.x < CNT_N) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int idx = blockIdx.x*blockDim.x + threadIdx.x;
  if (idx < n) {
    double point[3], dn[3], dscale = 1, power = 0;
    int medium = -1, seq_index = 0, n_isect = 0, skip_shell = -1;
    const RayState r = { point, dn, dscale, power, medium, seq_index, n_isect, skip_shell };
    fetch_ray<MC, FEAT_ALL>(p, idx, r);
    if (p.max_isect <= 0) {                                                      // ray.py:96-98 before the first segment
      atomicAdd(&s_cnt[CNT_DEPTH], 1u);
      finish_ray<MC>(p, idx, r, s_cnt);
    } else {
      pool_store(pool, idx, point, dn, power, dscale, idx, medium, seq_index);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) flush_counters(p, s_cnt);
}

// nearest intersection of every ray of the pool: hits[slot] = (t, face index | -1)
#ifndef ODW_WF_BLOCKS
#define ODW_WF_BLOCKS 4          // 64 registers, 32 warps per SM: the traversal is bound by node-fetch latency (measured 2, 3, 4, 5: 2.03, 2.09, 2.19, 2.03e9 segments/s)
#endif
#ifndef ODW_WF_FETCH_MIN
#define ODW_WF_FETCH_MIN 16       // measured 1 / 4 / 8 / 16 / 24 / 32: 3.41 / 3.44 / 3.48 / 3.53 / 3.47 / 2.90e9 segments/s on hugeArray
#endif
#ifndef ODW_WF_THREADS
#define ODW_WF_THREADS 1024     // one CTA per SM: ONE staged copy of the tree's top serves all 32 warps of the SM
#endif
template <int FEAT>
__global__ void __launch_bounds__(ODW_WF_THREADS, (ODW_WF_BLOCKS*256)/ODW_WF_THREADS) wf_traverse(const __grid_constant__ TraceParams p, WfPool pool, double2* hits,
                                                      unsigned int n, unsigned int* fetch_counter, const unsigned int* __restrict__ order,
                                                      WfPool ordered, int n_staged) {
  // the first n_staged nodes (breadth-first order: the top of the tree, all of it for hugeArray) live in shared memory: an inner
  // step then waits for a shared-memory read instead of an L1 / L2 round trip, which is what this kernel spends its time on
  extern __shared__ __align__(16) float4 s_nodes[];
  {
    const float4* g = reinterpret_cast<const float4*>(p.scene.bvh);
    for (int w = threadIdx.x; w < 4*n_staged; w += blockDim.x) s_nodes[staged_word(w >> 2, w & 3)] = __ldg(g + w);
    __syncthreads();
  }
  const unsigned int lane = threadIdx.x & 31u;
  bool have = false, exhausted = false;
  unsigned int slot = 0;
  double s[3] = {0, 0, 0}, dn[3] = {0, 0, 1};
  int medium = -1, seq_index = 0;
  BvhTrav tr;
  tr.cur = TRAV_DONE; tr.sp = 0;
  int stack_ref[ODW_BVH_STACK]; float stack_t[ODW_BVH_STACK];
  const BvhStack st = { stack_ref, stack_t };
  for (;;) {
    const unsigned int need = __ballot_sync(0xffffffffu, !have);
    // A fetch stalls the whole warp for a trip to the pool, so it waits until ODW_WF_FETCH_MIN lanes are free (or none has work)
    if (!exhausted && (__popc(need) >= ODW_WF_FETCH_MIN || need == 0xffffffffu)) {
      // lanes without a ray take the next pool slots: one atomic per warp
      const int leader = __ffs(need) - 1;
      unsigned int base = 0;
      if ((int)lane == leader) base = atomicAdd(fetch_counter, (unsigned int)__popc(need));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (!have) {
        slot = base + __popc(need & ((1u << lane) - 1u));
        if (slot < n) {
          const unsigned int src = order ? __ldg(order + slot) : slot;      // the k-th ray in coherence order
          const double2 a0 = pool.a0[src], a1 = pool.a1[src], a2 = pool.a2[src];
          const ulonglong2 a4 = pool.a4[src];
          if (order) {
            // the ray moves to slot k of the ordered pool (neighbouring lanes hold neighbouring k: coalesced stores), where the
            // interaction kernel finds it next to its result without chasing the order a second time
            ordered.a0[slot] = a0; ordered.a1[slot] = a1; ordered.a2[slot] = a2; ordered.a3[slot] = pool.a3[src]; ordered.a4[slot] = a4;
          }
          s[0] = a0.x; s[1] = a0.y; s[2] = a1.x; dn[0] = a1.y; dn[1] = a2.x; dn[2] = a2.y;
          medium = (int)(unsigned int)a4.y; seq_index = (int)(unsigned int)(a4.y >> 32);
          bvh_begin(tr, p, s, dn);
          have = true;
        }
      }
      exhausted = base + (unsigned int)__popc(need) >= n;
    }
    if (!__any_sync(0xffffffffu, have)) break;
    // descend: every lane walks inner nodes until it holds a leaf or has finished
    while (__any_sync(0xffffffffu, have && tr.cur >= 0)) {
      if (have && tr.cur >= 0) bvh_inner_step(tr, st, p, s_nodes, n_staged);
    }
    // leaves: exact face tests, all lanes that hold one together
    if (have && tr.cur < TRAV_DONE) bvh_leaf_step<FEAT>(tr, st, p, s, dn, medium, seq_index);
    if (have && tr.cur == TRAV_DONE) {
      const double tol = p.tol;
      double t = 0; int fi = -1;
      if (tr.h.fA >= 0) {                                                        // final choice, ray.py:438-452
        if (tr.h.fB >= 0 && tr.h.tB < tr.h.tA + 2*tol) { t = tr.h.tB; fi = tr.h.fB; } else { t = tr.h.tA; fi = tr.h.fA; }
      }
      hits[slot] = make_double2(t, __longlong_as_double((long long)fi));
      have = false;
    }
  }
}

// ---- 4-wide traversal with compact sphere leaves ------------------------------------------------------------------
// The binary traversal above is bound by instruction issue (72 % of the issue slots, profiles/r01_v15): ~25 inner steps of
// ~70 instructions and two or three 272-byte face records per ray.  This one halves the steps (four children per node, their
// boxes tested with one FFMA per plane: the near / far plane of each axis is picked by the sign of the direction when the
// node is read, t = plane * inv + (-origin * inv)), keeps the whole tree AND the primitives of sphere-array scenes in shared
// memory (hugeArray: 500 nodes + 1500 spheres = 112 KB), tests a whole sphere from a 32-byte record, and — for Monte-Carlo
// rays — draws the ray again from its number when it reaches the traversal in coherence order instead of gathering it from
// the pool (five 16-byte reads at unrelated addresses cost five 128-byte lines of DRAM: 641 B per ray measured).
struct Trav4 {
  NearestHit h;
  float ix, iy, iz, cx, cy, cz, limf;   // t(plane) = plane * i + c, c = -origin * i
  int wx, wy, wz;                       // word of the NEAR plane of each axis (far = near ^ 1)
  int cur, sp;
};
struct Stack4 { uint2* e; };            // entry = (key, child reference); key = entry distance (float bits, low 2 bits = child slot)

__device__ __forceinline__ int staged_word4(int node, int w) { return (node << 3) | (w ^ (node & 7)); }   // 16-byte words; swizzled like staged_word

__device__ __forceinline__ void bvh4_begin(Trav4& tr, const TraceParams& p, const double* s, const double* dn) {
  const double tmax = p.max_len + p.tol;
  tr.h.tA = 1e300; tr.h.tB = 1e300; tr.h.lim = tmax; tr.h.fA = -1; tr.h.fB = -1;
  // an axis-parallel direction gets a huge finite reciprocal: plane * inv - origin * inv stays finite (no inf - inf) and sorts
  // the ray inside / outside the slab like the exact test does, up to an offset the culling margin covers
  const float ix = fminf(fmaxf(__frcp_rn((float)dn[0]), -1e30f), 1e30f), iy = fminf(fmaxf(__frcp_rn((float)dn[1]), -1e30f), 1e30f),
              iz = fminf(fmaxf(__frcp_rn((float)dn[2]), -1e30f), 1e30f);
  tr.ix = ix; tr.iy = iy; tr.iz = iz;
  tr.cx = -(float)s[0]*ix; tr.cy = -(float)s[1]*iy; tr.cz = -(float)s[2]*iz;
  tr.wx = ix < 0.0f ? 1 : 0; tr.wy = iy < 0.0f ? 3 : 2; tr.wz = iz < 0.0f ? 5 : 4;
  tr.limf = (float)tmax*1.000002f;
  tr.cur = 0; tr.sp = 0;
}

// next stacked child that can still matter: ONE 8-byte load per entry, nothing else to chase
__device__ __forceinline__ int bvh4_pop(Trav4& tr, const Stack4& st) {
  while (tr.sp > 0) {
    --tr.sp;
    const uint2 e = st.e[tr.sp];
    if (__uint_as_float(e.x & ~3u) <= tr.limf) return (int)e.y;
  }
  return TRAV_DONE;
}

#define ODW_CAS(a, b) { const unsigned int lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }
__device__ __forceinline__ int ref_of(const int4 r, unsigned int key) {
  return (key & 2u) ? ((key & 1u) ? r.w : r.z) : ((key & 1u) ? r.y : r.x);
}
template <bool ALLSTAGED>
__device__ __forceinline__ void bvh4_inner_step(Trav4& tr, const Stack4& st, const TraceParams& p, const float4* s_nodes, int n_staged) {
  const int n = tr.cur;
  float4 nx, fx, ny, fy, nz, fz; int4 ref;
  if (ALLSTAGED || n < n_staged) {
    const int base = n << 3, sw = n & 7;
    nx = s_nodes[base | (tr.wx ^ sw)]; fx = s_nodes[base | (tr.wx ^ 1 ^ sw)];
    ny = s_nodes[base | (tr.wy ^ sw)]; fy = s_nodes[base | (tr.wy ^ 1 ^ sw)];
    nz = s_nodes[base | (tr.wz ^ sw)]; fz = s_nodes[base | (tr.wz ^ 1 ^ sw)];
    const float4 rr = s_nodes[base | (6 ^ sw)];
    ref = make_int4(__float_as_int(rr.x), __float_as_int(rr.y), __float_as_int(rr.z), __float_as_int(rr.w));
  } else {
    const float4* q = reinterpret_cast<const float4*>(p.scene.bvh4 + n);
    nx = __ldg(q + tr.wx); fx = __ldg(q + (tr.wx ^ 1)); ny = __ldg(q + tr.wy); fy = __ldg(q + (tr.wy ^ 1));
    nz = __ldg(q + tr.wz); fz = __ldg(q + (tr.wz ^ 1));
    ref = __ldg(reinterpret_cast<const int4*>(q + 6));
  }
  const float ix = tr.ix, iy = tr.iy, iz = tr.iz, cx = tr.cx, cy = tr.cy, cz = tr.cz, limf = tr.limf;
  // entry / exit distance of each child box, clamped to [0, limf]: hit <=> entry <= exit
  const float a0 = fmaxf(fmaxf(fmaf(nx.x, ix, cx), fmaf(ny.x, iy, cy)), fmaxf(fmaf(nz.x, iz, cz), 0.0f));
  const float b0 = fminf(fminf(fmaf(fx.x, ix, cx), fmaf(fy.x, iy, cy)), fminf(fmaf(fz.x, iz, cz), limf));
  const float a1 = fmaxf(fmaxf(fmaf(nx.y, ix, cx), fmaf(ny.y, iy, cy)), fmaxf(fmaf(nz.y, iz, cz), 0.0f));
  const float b1 = fminf(fminf(fmaf(fx.y, ix, cx), fmaf(fy.y, iy, cy)), fminf(fmaf(fz.y, iz, cz), limf));
  const float a2 = fmaxf(fmaxf(fmaf(nx.z, ix, cx), fmaf(ny.z, iy, cy)), fmaxf(fmaf(nz.z, iz, cz), 0.0f));
  const float b2 = fminf(fminf(fmaf(fx.z, ix, cx), fmaf(fy.z, iy, cy)), fminf(fmaf(fz.z, iz, cz), limf));
  const float a3 = fmaxf(fmaxf(fmaf(nx.w, ix, cx), fmaf(ny.w, iy, cy)), fmaxf(fmaf(nz.w, iz, cz), 0.0f));
  const float b3 = fminf(fminf(fmaf(fx.w, ix, cx), fmaf(fy.w, iy, cy)), fminf(fmaf(fz.w, iz, cz), limf));
  // sort the (up to four) hit children by entry distance: non-negative floats order like their bit patterns, the child slot
  // rides in the two lowest mantissa bits (the distance only loses precision downwards: conservative for the pruning at pop)
  unsigned int k0 = a0 <= b0 ? (__float_as_uint(a0) & ~3u) : 0xffffffffu;
  unsigned int k1 = a1 <= b1 ? ((__float_as_uint(a1) & ~3u) | 1u) : 0xffffffffu;
  unsigned int k2 = a2 <= b2 ? ((__float_as_uint(a2) & ~3u) | 2u) : 0xffffffffu;
  unsigned int k3 = a3 <= b3 ? ((__float_as_uint(a3) & ~3u) | 3u) : 0xffffffffu;
  ODW_CAS(k0, k1) ODW_CAS(k2, k3) ODW_CAS(k0, k2) ODW_CAS(k1, k3) ODW_CAS(k1, k2)
  if (k0 == 0xffffffffu) { tr.cur = bvh4_pop(tr, st); return; }
  // farthest first, so that the nearest of the stacked children is popped first
  if (k1 != 0xffffffffu) {
    if (k2 != 0xffffffffu) {
      if (k3 != 0xffffffffu && tr.sp < ODW_BVH_STACK) { st.e[tr.sp] = make_uint2(k3, (unsigned int)ref_of(ref, k3)); ++tr.sp; }
      if (tr.sp < ODW_BVH_STACK) { st.e[tr.sp] = make_uint2(k2, (unsigned int)ref_of(ref, k2)); ++tr.sp; }
    }
    if (tr.sp < ODW_BVH_STACK) { st.e[tr.sp] = make_uint2(k1, (unsigned int)ref_of(ref, k1)); ++tr.sp; }
  }
  tr.cur = ref_of(ref, k0);
}

// whole sphere from its compact record: the arithmetic of test_face's sphere branch with a = 1 and no axial window
__device__ __forceinline__ void test_sphere(const DSphere& sp, const int2 info, const TraceParams& p, const double* s, const double* dn,
                                            int medium, int seq_index, bool filter, NearestHit& h) {
  if (filter) {        // launch-uniform: sequential mode or a non-empty ignore list
    if (p.sequential) {
      if (seq_index >= 128) return;
      const ulonglong2 m = __ldg(p.scene.group_seqmask + info.y);
      if (!(((seq_index < 64 ? m.x : m.y) >> (seq_index & 63)) & 1ull)) return;
    }
    if (info.y < 256 && ((p.ignore_mask[info.y >> 6] >> (info.y & 63)) & 1ull)) return;
  }
  const double tol = p.tol, limit = h.lim;
  const double w0 = s[0]-sp.cx, w1 = s[1]-sp.cy, w2 = s[2]-sp.cz;
  const double b = dot3(w0, w1, w2, dn), c = dot3(w0, w1, w2, w0, w1, w2) - sp.r*sp.r;
  const double disc = b*b - c;
  if (!(disc >= 0)) return;
  const double sq = fast_sqrt(disc);
  const double q = -(b + (b >= 0 ? sq : -sq));
  const double r1 = (q != 0) ? c*fast_rcp(q) : 0.0;
  const double tn = fmin(q, r1), tf = fmax(q, r1);
  if (tn > tol && tn < limit) accept_hit(tn, info.x, info.y, medium, tol, h);
  if (tf > tol && tf < limit) accept_hit(tf, info.x, info.y, medium, tol, h);
}

template <int FEAT>
__device__ __forceinline__ void bvh4_leaf_step(Trav4& tr, const Stack4& st, const TraceParams& p,
                                               const int32_t* prims, const DSphere* spheres, const int2* sphere_info, bool filter,
                                               const double* s, const double* dn, int medium, int seq_index) {
  const int v = -2 - tr.cur, first = v >> 3, count = (v & 7) + 1;
  const double tmax = p.max_len + p.tol;
  for (int k = 0; k < count; ++k) {
    const int prim = prims[first + k];
    if (prim < 0) test_sphere(spheres[~prim], sphere_info[~prim], p, s, dn, medium, seq_index, filter, tr.h);
    else test_face<true, FEAT>(p.scene.faces[prim], prim, p, s, dn, medium, seq_index, tmax, tr.h);
  }
  tr.limf = (float)tr.h.lim*1.000002f;
  tr.cur = bvh4_pop(tr, st);
}

// nearest intersection of every ray of the pool, 4-wide tree.  REGEN: Monte-Carlo rays of the first bounce are drawn again from
// their number (order[slot]) instead of being read from the pool; the traversal writes them, in coherence order, into `ordered`.
#ifndef ODW_WF4_THREADS
#define ODW_WF4_THREADS 768        // 80 registers: measured 1024 / 768 / 512 threads: 4.50 / 4.60 / 4.49e9 segments/s on hugeArray
#endif
template <int FEAT, bool REGEN, bool ALLSTAGED>
__global__ void __launch_bounds__(ODW_WF4_THREADS, 1) wf_traverse4(const __grid_constant__ TraceParams p, WfPool pool, double2* hits,
                                                      unsigned int n, unsigned int* fetch_counter, const unsigned int* __restrict__ order,
                                                      WfPool ordered, int n_staged, int stage_prims) {
  extern __shared__ __align__(16) float4 s_nodes[];
  // shared memory: n_staged nodes, then (stage_prims) the compact spheres, their (face, group) pairs and the leaf entries
  DSphere* s_spheres = reinterpret_cast<DSphere*>(s_nodes + 8*(size_t)n_staged);
  int2* s_info = reinterpret_cast<int2*>(s_spheres + (stage_prims ? p.scene.n_spheres : 0));
  int32_t* s_prims = reinterpret_cast<int32_t*>(s_info + (stage_prims ? p.scene.n_spheres : 0));
  {
    const float4* g = reinterpret_cast<const float4*>(p.scene.bvh4);
    for (int w = threadIdx.x; w < 8*n_staged; w += blockDim.x) s_nodes[staged_word4(w >> 3, w & 7)] = __ldg(g + w);
    if (stage_prims) {
      const double2* gs = reinterpret_cast<const double2*>(p.scene.spheres);
      for (int k = threadIdx.x; k < p.scene.n_spheres; k += blockDim.x) {
        const double2 v0 = __ldg(gs + 2*k), v1 = __ldg(gs + 2*k + 1);
        s_spheres[k].cx = v0.x; s_spheres[k].cy = v0.y; s_spheres[k].cz = v1.x; s_spheres[k].r = v1.y;
        s_info[k] = __ldg(p.scene.sphere_info + k);
      }
      for (int k = threadIdx.x; k < p.scene.n_bvh4_prims; k += blockDim.x) s_prims[k] = __ldg(p.scene.bvh4_prims + k);
    }
    __syncthreads();
  }
  // ALLSTAGED instances are launched only when the compact primitives are staged too: their pointers are known to be shared
  // memory at compile time (LDS), the others are generic (shared or global)
  const DSphere* spheres = (ALLSTAGED || stage_prims) ? s_spheres : p.scene.spheres;
  const int2* sphere_info = (ALLSTAGED || stage_prims) ? s_info : p.scene.sphere_info;
  const int32_t* prims = (ALLSTAGED || stage_prims) ? s_prims : p.scene.bvh4_prims;
  const bool filter = p.sequential || (p.ignore_mask[0] | p.ignore_mask[1] | p.ignore_mask[2] | p.ignore_mask[3]) != 0ull;
  const unsigned int lane = threadIdx.x & 31u;
  bool have = false, exhausted = false;
  unsigned int slot = 0;
  double s[3] = {0, 0, 0}, dn[3] = {0, 0, 1};
  int medium = -1, seq_index = 0;
  Trav4 tr;
  tr.cur = TRAV_DONE; tr.sp = 0;
  uint2 stack_e[ODW_BVH_STACK];
  const Stack4 st = { stack_e };
  for (;;) {
    const unsigned int need = __ballot_sync(0xffffffffu, !have);
    if (!exhausted && (__popc(need) >= ODW_WF_FETCH_MIN || need == 0xffffffffu)) {
      const int leader = __ffs(need) - 1;
      unsigned int base = 0;
      if ((int)lane == leader) base = atomicAdd(fetch_counter, (unsigned int)__popc(need));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (!have) {
        slot = base + __popc(need & ((1u << lane) - 1u));
        if (slot < n) {
          const unsigned int src = order ? __ldg(order + slot) : slot;      // the k-th ray in coherence order
          if (REGEN) {
            // Monte-Carlo rays are a function of their number: draw ray `src` again and put it at slot k of the ordered pool
            double dscale = 1, power = 0;
            int n_isect = 0, skip_shell = -1;
            const RayState r = { s, dn, dscale, power, medium, seq_index, n_isect, skip_shell };
            fetch_ray<true, FEAT_ALL>(p, src, r);
            ordered.a0[slot] = make_double2(s[0], s[1]); ordered.a1[slot] = make_double2(s[2], dn[0]); ordered.a2[slot] = make_double2(dn[1], dn[2]);
            ordered.a3[slot] = make_double2(power, dscale);
            ordered.a4[slot] = make_ulonglong2((unsigned long long)src, (unsigned long long)(unsigned int)medium | ((unsigned long long)(unsigned int)seq_index << 32));
          } else {
            const double2 a0 = pool.a0[src], a1 = pool.a1[src], a2 = pool.a2[src];
            const ulonglong2 a4 = pool.a4[src];
            if (order) { ordered.a0[slot] = a0; ordered.a1[slot] = a1; ordered.a2[slot] = a2; ordered.a3[slot] = pool.a3[src]; ordered.a4[slot] = a4; }
            s[0] = a0.x; s[1] = a0.y; s[2] = a1.x; dn[0] = a1.y; dn[1] = a2.x; dn[2] = a2.y;
            medium = (int)(unsigned int)a4.y; seq_index = (int)(unsigned int)(a4.y >> 32);
          }
          bvh4_begin(tr, p, s, dn);
          have = true;
        }
      }
      exhausted = base + (unsigned int)__popc(need) >= n;
    }
    if (!__any_sync(0xffffffffu, have)) break;
    while (__any_sync(0xffffffffu, have && tr.cur >= 0)) {
      if (have && tr.cur >= 0) bvh4_inner_step<ALLSTAGED>(tr, st, p, s_nodes, n_staged);
    }
    if (have && tr.cur < TRAV_DONE) bvh4_leaf_step<FEAT>(tr, st, p, prims, spheres, sphere_info, filter, s, dn, medium, seq_index);
    if (have && tr.cur == TRAV_DONE) {
      const double tol = p.tol;
      double t = 0; int fi = -1;
      if (tr.h.fA >= 0) {                                                        // final choice, ray.py:438-452
        if (tr.h.fB >= 0 && tr.h.tB < tr.h.tA + 2*tol) { t = tr.h.tB; fi = tr.h.fB; } else { t = tr.h.tA; fi = tr.h.fA; }
      }
      hits[slot] = make_double2(t, __longlong_as_double((long long)fi));
      have = false;
    }
  }
}

// keys only: Monte-Carlo waves whose first traversal draws the rays again in coherence order (wf_traverse4<.., true>)
__global__ void __launch_bounds__(256) wf_generate_keys(const __grid_constant__ TraceParams p, WfPool pool, unsigned int n) {
  const unsigned int idx = blockIdx.x*blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double point[3], dn[3], dscale = 1, power = 0;
  if (p.src.kind == ODW_SRC_POINT_SPHERICAL && p.src.focal == 0.0) {
    // every ray starts at the same point: the key is the direction cell only, and a rotation keeps neighbours neighbours, so the
    // direction in the SOURCE frame, in fp32, orders the rays as well as the world direction would (the key decides nothing
    // but the order in which the rays are traversed)
    double u0, u1, first, phi;
    philox_uniform2(p.seed, (uint32_t)p.src.source_id, p.first_ray + idx, 0u, u0, u1);
    sample_source(p.src, u0, u1, first, phi);
    float st, ct, sp, cp;
    __sincosf((float)first, &st, &ct); __sincosf((float)phi, &sp, &cp);
    point[0] = point[1] = point[2] = 0.0;
    dn[0] = st*sp; dn[1] = -st*cp; dn[2] = ct;
    pool.key[idx] = ray_sort_key(point, dn, pool.bound);
    return;
  }
  int medium = -1, seq_index = 0, n_isect = 0, skip_shell = -1;
  const RayState r = { point, dn, dscale, power, medium, seq_index, n_isect, skip_shell };
  fetch_ray<true, FEAT_ALL>(p, idx, r);
  pool.key[idx] = ray_sort_key(point, dn, pool.bound);
}

// surface interaction of every ray of pool_in with its hit; survivors are appended to pool_out
// KEYS: the survivors' coherence keys are wanted (only when the NEXT bounce is sorted too; by default only the first one is)
template <bool MC, bool KEYS>
__global__ void __launch_bounds__(256) wf_interact(const __grid_constant__ TraceParams p, WfPool pool_in, const double2* __restrict__ hits,
                                                   WfPool pool_out, unsigned int n, unsigned int* n_next, int bounce) {
  __shared__ unsigned int s_cnt[CNT_N];
  if (threadIdx.x < CNT_N) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int idx = blockIdx.x*blockDim.x + threadIdx.x;
  bool survive = false;
  double point[3], dn[3], dscale = 1, power = 0;
  int medium = -1, seq_index = 0, n_isect = bounce + 1, skip_shell = -1;   // every ray of this wave has done `bounce` segments before
  unsigned long long i = 0;
  if (idx < n) {
    const double2 a0 = pool_in.a0[idx], a1 = pool_in.a1[idx], a2 = pool_in.a2[idx], a3 = pool_in.a3[idx];
    const ulonglong2 a4 = pool_in.a4[idx];
    const double2 h = hits[idx];
    point[0] = a0.x; point[1] = a0.y; point[2] = a1.x; dn[0] = a1.y; dn[1] = a2.x; dn[2] = a2.y;
    power = a3.x; dscale = a3.y; i = a4.x;
    medium = (int)(unsigned int)a4.y; seq_index = (int)(unsigned int)(a4.y >> 32);
    const RayState r = { point, dn, dscale, power, medium, seq_index, n_isect, skip_shell };
    bool done = interact<MC, FEAT_ALL>(p, p.scene.faces, nullptr, p.scene.groups, (int)__double_as_longlong(h.y), h.x, i, r, s_cnt);
    if (!done && n_isect >= p.max_isect) { atomicAdd(&s_cnt[CNT_DEPTH], 1u); done = true; }   // ray.py:96-98
    if (done) finish_ray<MC>(p, i, r, s_cnt);
    survive = !done;
  }
  // compaction: one atomic per warp (warp ballot + prefix count)
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int alive = __ballot_sync(0xffffffffu, survive);
  if (alive) {
    const int leader = __ffs(alive) - 1;
    unsigned int base = 0;
    if ((int)lane == leader) base = atomicAdd(n_next, (unsigned int)__popc(alive));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (survive) pool_store<KEYS>(pool_out, base + __popc(alive & ((1u << lane) - 1u)), point, dn, power, dscale, i, medium, seq_index);
  }
  __syncthreads();
  if (threadIdx.x == 0) flush_counters(p, s_cnt);
}

// the last survivors of a wave: one lane per ray runs the rest of its bounce loop
template <bool MC>
__global__ void __launch_bounds__(256) wf_tail(const __grid_constant__ TraceParams p, WfPool pool, unsigned int n, int bounce) {
  __shared__ unsigned int s_cnt[CNT_N];
  if (threadIdx.x < CNT_N) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int idx = blockIdx.x*blockDim.x + threadIdx.x;
  if (idx < n) {
    double point[3], dn[3], dscale, power;
    int medium, seq_index, n_isect = bounce, skip_shell = -1;
    const double2 a0 = pool.a0[idx], a1 = pool.a1[idx], a2 = pool.a2[idx], a3 = pool.a3[idx];
    const ulonglong2 a4 = pool.a4[idx];
    point[0] = a0.x; point[1] = a0.y; point[2] = a1.x; dn[0] = a1.y; dn[1] = a2.x; dn[2] = a2.y;
    power = a3.x; dscale = a3.y;
    const unsigned long long i = a4.x;
    medium = (int)(unsigned int)a4.y; seq_index = (int)(unsigned int)(a4.y >> 32);
    const RayState r = { point, dn, dscale, power, medium, seq_index, n_isect, skip_shell };
    for (;;) {
      if (n_isect >= p.max_isect) { atomicAdd(&s_cnt[CNT_DEPTH], 1u); break; }
      ++n_isect;
      double t;
      const int fi = find_nearest_bvh<FEAT_ALL>(p, point, dn, medium, seq_index, p.max_len, t);
      if (interact<MC, FEAT_ALL>(p, p.scene.faces, nullptr, p.scene.groups, fi, t, i, r, s_cnt)) break;
    }
    finish_ray<MC>(p, i, r, s_cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) flush_counters(p, s_cnt);
}

// ---- launch helpers used by odw_api.cu ---------------------------------------------------------
extern "C" size_t odw_wf_pool_bytes_per_ray(void) { return 4*sizeof(double2) + sizeof(ulonglong2) + sizeof(unsigned int); }

static WfPool make_pool(void* base, size_t cap, float bound = 1.0f) {
  WfPool pl;
  char* b = static_cast<char*>(base);
  pl.a0 = reinterpret_cast<double2*>(b); b += cap*sizeof(double2);
  pl.a1 = reinterpret_cast<double2*>(b); b += cap*sizeof(double2);
  pl.a2 = reinterpret_cast<double2*>(b); b += cap*sizeof(double2);
  pl.a3 = reinterpret_cast<double2*>(b); b += cap*sizeof(double2);
  pl.a4 = reinterpret_cast<ulonglong2*>(b); b += cap*sizeof(ulonglong2);
  pl.key = reinterpret_cast<unsigned int*>(b);
  pl.bound = bound;
  return pl;
}

extern "C" cudaError_t odw_wf_generate(const TraceParams* p, bool mc, void* pool, size_t cap, float bound, unsigned int n, cudaStream_t st) {
  const WfPool pl = make_pool(pool, cap, bound);
  const unsigned int blocks = (n + 255u)/256u;
  if (mc) wf_generate<true><<<blocks, 256, 0, st>>>(*p, pl, n); else wf_generate<false><<<blocks, 256, 0, st>>>(*p, pl, n);
  return cudaGetLastError();
}

// nodes of the tree that fit the shared memory one CTA may use (the rest is read through L1 / L2)
static int wf_staged_nodes(int n_nodes) {
  // shared memory and L1 share 256 KB per SM: 96 KB of nodes (1536, all of hugeArray's 1499) leave the L1 enough room for the
  // traversal stacks and the face records
  const int per_cta = ODW_WF_THREADS >= 1024 ? 96*1024 : (ODW_WF_THREADS >= 512 ? 96*1024 : 48*1024);
  int n = per_cta/(int)sizeof(BvhNode2);
  n = n < n_nodes ? n : n_nodes;
  return n & ~1;                       // whole rows of the swizzle (two nodes per 128-byte row)
}

extern "C" cudaError_t odw_wf_traverse(const TraceParams* p, void* pool, size_t cap, void* hits, unsigned int n,
                                       unsigned int* fetch_counter, const unsigned int* order, void* pool_ordered, int need, int blocks, cudaStream_t st) {
  const unsigned int want = (n + ODW_WF_THREADS - 1u)/ODW_WF_THREADS, grid = (unsigned int)blocks < want ? (unsigned int)blocks : want;
  const WfPool pl = make_pool(pool, cap), po = make_pool(order ? pool_ordered : pool, cap);
  const int n_staged = wf_staged_nodes(p->scene.n_bvh_nodes);
  const size_t smem = (size_t)n_staged*sizeof(BvhNode2);
  // need: FEAT_* bits of the launch; only FEAT_EXT (even-asphere faces) concerns the traversal
  if (need & FEAT_EXT) {
    cudaFuncSetAttribute(wf_traverse<FEAT_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wf_traverse<FEAT_ALL><<<grid, ODW_WF_THREADS, smem, st>>>(*p, pl, static_cast<double2*>(hits), n, fetch_counter, order, po, n_staged);
  } else {
    cudaFuncSetAttribute(wf_traverse<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wf_traverse<0><<<grid, ODW_WF_THREADS, smem, st>>>(*p, pl, static_cast<double2*>(hits), n, fetch_counter, order, po, n_staged);
  }
  return cudaGetLastError();
}

// shared memory of wf_traverse4: as many nodes as fit, and the compact primitives when they fit next to them
static void wf4_layout(const DScene& sc, int* n_staged, int* stage_prims, size_t* smem) {
  const size_t budget = 200*1024;                                   // of the 227 KB a CTA may have; one CTA per SM
  const size_t prim_bytes = (size_t)sc.n_spheres*(sizeof(DSphere) + sizeof(int2)) + (size_t)sc.n_bvh4_prims*sizeof(int32_t);
  *stage_prims = (sc.n_spheres > 0 && prim_bytes <= 96*1024) ? 1 : 0;
  const size_t left = budget - (*stage_prims ? ((prim_bytes + 15) & ~(size_t)15) : 0);
  int n = (int)std::min<size_t>((size_t)sc.n_bvh4_nodes, left/sizeof(Bvh4Node));
  *n_staged = n;
  *smem = (size_t)n*sizeof(Bvh4Node) + (*stage_prims ? ((prim_bytes + 15) & ~(size_t)15) : 0) + 16;
}

extern "C" cudaError_t odw_wf_traverse4(const TraceParams* p, void* pool, size_t cap, void* hits, unsigned int n,
                                        unsigned int* fetch_counter, const unsigned int* order, void* pool_ordered, int need, int regen,
                                        int blocks, cudaStream_t st) {
  const unsigned int want = (n + ODW_WF4_THREADS - 1u)/ODW_WF4_THREADS, grid = (unsigned int)blocks < want ? (unsigned int)blocks : want;
  const WfPool pl = make_pool(pool, cap), po = make_pool(order ? pool_ordered : pool, cap);
  int n_staged, stage_prims; size_t smem;
  wf4_layout(p->scene, &n_staged, &stage_prims, &smem);
#define ODW_LAUNCH4(F, R, A) { cudaFuncSetAttribute(wf_traverse4<F, R, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    wf_traverse4<F, R, A><<<grid, ODW_WF4_THREADS, smem, st>>>(*p, pl, static_cast<double2*>(hits), n, fetch_counter, order, po, n_staged, stage_prims); }
  const bool all = n_staged == p->scene.n_bvh4_nodes && stage_prims;
  if (need & FEAT_EXT) { if (regen) ODW_LAUNCH4(FEAT_ALL, true, false) else ODW_LAUNCH4(FEAT_ALL, false, false) }
  else if (all) { if (regen) ODW_LAUNCH4(0, true, true) else ODW_LAUNCH4(0, false, true) }
  else { if (regen) ODW_LAUNCH4(0, true, false) else ODW_LAUNCH4(0, false, false) }
#undef ODW_LAUNCH4
  return cudaGetLastError();
}

extern "C" cudaError_t odw_wf_generate_keys(const TraceParams* p, void* pool, size_t cap, float bound, unsigned int n, cudaStream_t st) {
  wf_generate_keys<<<(n + 255u)/256u, 256, 0, st>>>(*p, make_pool(pool, cap, bound), n);
  return cudaGetLastError();
}

extern "C" cudaError_t odw_wf_interact(const TraceParams* p, bool mc, void* pool_in, void* hits, void* pool_out, size_t cap, float bound,
                                       unsigned int n, unsigned int* n_next, int bounce, int keys, cudaStream_t st) {
  const unsigned int blocks = (n + 255u)/256u;
  const WfPool pi = make_pool(pool_in, cap, bound), po = make_pool(pool_out, cap, bound);
  const double2* h = static_cast<const double2*>(hits);
  if (mc) { if (keys) wf_interact<true, true><<<blocks, 256, 0, st>>>(*p, pi, h, po, n, n_next, bounce); else wf_interact<true, false><<<blocks, 256, 0, st>>>(*p, pi, h, po, n, n_next, bounce); }
  else { if (keys) wf_interact<false, true><<<blocks, 256, 0, st>>>(*p, pi, h, po, n, n_next, bounce); else wf_interact<false, false><<<blocks, 256, 0, st>>>(*p, pi, h, po, n, n_next, bounce); }
  return cudaGetLastError();
}

extern "C" cudaError_t odw_wf_tail(const TraceParams* p, bool mc, void* pool, size_t cap, unsigned int n, int bounce, cudaStream_t st) {
  const unsigned int blocks = (n + 255u)/256u;
  if (mc) wf_tail<true><<<blocks, 256, 0, st>>>(*p, make_pool(pool, cap), n, bounce);
  else wf_tail<false><<<blocks, 256, 0, st>>>(*p, make_pool(pool, cap), n, bounce);
  return cudaGetLastError();
}

// coherence order of the first n rays of a pool: order[k] = pool slot of the k-th ray by key (radix sort of (key, slot) pairs)
__global__ void wf_iota(unsigned int* v, unsigned int n) {
  const unsigned int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
extern "C" cudaError_t odw_wf_iota(unsigned int* v, unsigned int n, cudaStream_t st) {
  wf_iota<<<(n + 255u)/256u, 256, 0, st>>>(v, n);
  return cudaGetLastError();
}
// temp == nullptr: size query
extern "C" cudaError_t odw_wf_sort(void* temp, size_t* temp_bytes, void* pool, size_t cap, unsigned int* keys_out, const unsigned int* iota,
                                   unsigned int* order, unsigned int n, cudaStream_t st) {
  const WfPool pl = make_pool(pool, cap);
  return cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, (const unsigned int*)pl.key, keys_out, iota, order, (int)n, 0, ODW_SORT_KEY_BITS, st);
}

// (key, value) pairs by bits [begin_bit, end_bit) of the key; temp == nullptr: size query
extern "C" cudaError_t odw_sort_pairs(void* temp, size_t* temp_bytes, const unsigned int* keys_in, unsigned int* keys_out,
                                      const unsigned int* vals_in, unsigned int* vals_out, unsigned int n, int begin_bit, int end_bit, cudaStream_t st) {
  return cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit, st);
}

extern "C" int odw_wf_traverse_occupancy(int n_nodes) {
  int nb = 0;
  const size_t smem = (size_t)wf_staged_nodes(n_nodes)*sizeof(BvhNode2);
  cudaFuncSetAttribute(wf_traverse<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wf_traverse<0>, ODW_WF_THREADS, smem);
  return nb;
}
