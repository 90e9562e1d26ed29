// odw_kernels.cu — sm_100a trace kernels.
//
// One thread owns one ray for its whole life: the bounce loop of Ray.traceRay (reference
// freecad_elements/ray.py:36-281) runs in registers, so the only HBM traffic is the hit append (and the
// explicit-ray input when there is one).  MC rays are generated in the kernel from a Philox4x32-10
// counter = global ray index, so the result does not depend on how a ray range is split
// over launches or GPUs.  Small scenes are staged in shared memory and tested face by face (uniform loop,
// no divergence between lanes of a warp); large scenes walk a BVH with a per-thread short stack.
#ifndef ODW_BLOCK_SYNC
#define ODW_BLOCK_SYNC 0        // 1: re-converge the whole CTA once per bounce (measured slower since the fp32 culls shrank the loop)
#endif
#include "odw_trace.cuh"

#ifndef ODW_MIN_BLOCKS
#define ODW_MIN_BLOCKS 3          // 3 CTAs x 8 warps per SM (80 registers): the kernel is latency-bound, 24 warps beat 16 despite spills
#endif
#ifndef ODW_RAY_CHUNK
#define ODW_RAY_CHUNK 32        // rays a warp claims per atomic
#endif
#ifndef ODW_THREADS
#define ODW_THREADS 256          // threads per CTA of the trace kernel
#endif
template <bool MC, bool BVH, int FEAT>
__global__ void __launch_bounds__(ODW_THREADS, ODW_MIN_BLOCKS) trace_kernel(const __grid_constant__ TraceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DShell* sshells = reinterpret_cast<DShell*>(smem_raw);
  DFace* sfaces = reinterpret_cast<DFace*>(smem_raw + (size_t)p.scene.n_shells*sizeof(DShell));
  __shared__ unsigned int s_cnt[CNT_N];
  if (threadIdx.x < CNT_N) s_cnt[threadIdx.x] = 0;
  if (BVH) __syncthreads();
  if (!BVH) {
    // stage the scene (shells, then faces): 16-byte vector copies, coalesced
    const int4* src = reinterpret_cast<const int4*>(p.scene.shells);
    int4* dst = reinterpret_cast<int4*>(sshells);
    int n16 = (int)((size_t)p.scene.n_shells*sizeof(DShell)/16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    src = reinterpret_cast<const int4*>(p.scene.faces);
    dst = reinterpret_cast<int4*>(sfaces);
    n16 = (int)((size_t)p.scene.n_faces*sizeof(DFace)/16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    // widen the fp32 shell boxes by the culling margin; shells of ignored groups (IgnoredOpticalElements,
    // generic_source.py:23-37 / find.py:79-104) are moved out of reach (a point at 3e38) and lose their faces
    for (int i = threadIdx.x; i < p.scene.n_shells; i += blockDim.x) {
      DShell& sh = sshells[i];
      const bool ignored = sh.group < 256 && ((p.ignore_mask[sh.group >> 6] >> (sh.group & 63)) & 1ull);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        sh.lo[k] = ignored ? 3.0e38f : __fsub_rd(sh.lo[k], p.cull_margin);
        sh.hi[k] = ignored ? 3.0e38f : __fadd_ru(sh.hi[k], p.cull_margin);
      }
      if (ignored) sh.face_count = 0;
    }
    __syncthreads();
  }
  unsigned long long t_start = 0, c_start = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start)); c_start = clock64(); }
  const DGroup* __restrict__ groups = p.scene.groups;
  // Persistent lanes: a lane whose ray has ended takes its next ray and initialises it; then every lane with a live ray
  // does ONE bounce.  The warp re-converges once per bounce (the __any_sync below), so the expensive part (find nearest
  // intersection) always runs with all live lanes together, no matter how differently long the rays of a warp are.
  // Rays are claimed dynamically: a warp takes ODW_RAY_CHUNK consecutive ray numbers from the launch's counter (one
  // atomic per chunk) and hands them to its lanes as they become free.  With a fixed lane -> ray stride a 2^18-ray launch
  // gave each lane two or three rays of one to seven segments each, and the launch lasted as long as its unluckiest lane.
  // (Monte-Carlo rays are a function of their number alone, so who traces which ray does not change a result.)
  // (Measured and rejected for the BVH path, where only ~6 of 32 lanes are active on hugeArray because traversal lengths
  // differ wildly: alternating "traversal step" / "interaction" phases gated by the number of lanes still traversing.
  // The interaction + ray-initialisation code then runs once per few traversal steps instead of once per bounce and
  // costs more than the idle lanes it saves: 7.5e8 vs 9.3e8 segments/s.)
  const unsigned int lane = threadIdx.x & 31u;
  unsigned long long pool_next = 0;          // warp-uniform: next ray number of the warp's chunk
  unsigned int pool_left = 0;                // warp-uniform: rays left in it
  bool exhausted = false;                    // warp-uniform: the launch has no unclaimed rays
  unsigned long long i = 0;                  // number of the lane's current ray inside the launch
  double point[3] = {0, 0, 0}, dn[3] = {0, 0, 1}, dscale = 1, power = 0;
  int medium = -1, seq_index = 0, n_isect = 0, skip_shell = -1;
  const RayState r = { point, dn, dscale, power, medium, seq_index, n_isect, skip_shell };
  bool alive = false;
  for (;;) {
    const unsigned int need = __ballot_sync(0xffffffffu, !alive);
    if (need && !exhausted) {
      if (pool_left == 0) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(p.ray_counter, (unsigned long long)ODW_RAY_CHUNK);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= p.n_rays) exhausted = true;
        else { pool_next = b; pool_left = (unsigned int)min((unsigned long long)ODW_RAY_CHUNK, p.n_rays - b); }
      }
      if (pool_left) {
        const unsigned int k = __popc(need & ((1u << lane) - 1u));
        if (!alive && k < pool_left) {
          i = pool_next + k;
          if (!MC && p.in_order) i = __ldg(p.in_order + i);
          fetch_ray<MC, FEAT>(p, i, r); alive = true;
        }
        const unsigned int taken = min((unsigned int)__popc(need), pool_left);
        pool_next += taken; pool_left -= taken;
      }
    }
#if ODW_BLOCK_SYNC
    // block-wide re-convergence: all warps of a CTA stay in the same phase of the loop, so the CTA's instruction
    // working set is one phase (init / intersect / interact) instead of all of them at once
    if (!__syncthreads_or(alive || !exhausted)) break;
#else
    if (!__any_sync(0xffffffffu, alive)) { if (exhausted) break; else continue; }
#endif
    if (alive) {
      bool done;
      if (n_isect >= p.max_isect) { atomicAdd(&s_cnt[CNT_DEPTH], 1u); done = true; }      // ray.py:96-98
      else {
        ++n_isect;
        double t;
        const int fi = BVH ? find_nearest_bvh<FEAT>(p, point, dn, medium, seq_index, p.max_len, t)
                           : find_nearest_smem<FEAT>(sshells, sfaces, p, point, dn, medium, seq_index, skip_shell, p.max_len, t);
        done = interact<MC, FEAT>(p, BVH ? p.scene.faces : sfaces, BVH ? nullptr : sshells, groups, fi, t, i, r, s_cnt);
      }
      if (done) { finish_ray<MC>(p, i, r, s_cnt); alive = false; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    p.counters->dbg_cycles = clock64() - c_start; p.counters->dbg_ns = t_end - t_start;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&p.counters->segments, (unsigned long long)s_cnt[CNT_SEGMENTS]);
    if (!p.store_hits) atomicAdd(&p.counters->hits, (unsigned long long)s_cnt[CNT_HITS]);
    if (s_cnt[CNT_DROPPED]) atomicAdd(&p.counters->hits_dropped, (unsigned long long)s_cnt[CNT_DROPPED]);
    if (s_cnt[CNT_ESCAPED]) atomicAdd(&p.counters->escaped, (unsigned long long)s_cnt[CNT_ESCAPED]);
    if (s_cnt[CNT_DEPTH]) atomicAdd(&p.counters->depth_terminated, (unsigned long long)s_cnt[CNT_DEPTH]);
  }
}

// draws only (odw_sample_mc): same Philox counters and tables as the MC trace kernel.
// Surface sources: the three ways a point is drawn (triangle without rejection, analytic surface with natural bounds,
// trimmed face with the loop test) differ by an order of magnitude in work, and a warp that holds all three runs them one
// after the other.  Each WARP therefore takes ODW_SAMPLE_SPAN consecutive rays, picks their faces first, orders the rays by
// face class in its slice of shared memory and only then draws, 32 at a time: the lanes are in the same code except at the
// class boundaries.  The results are stored by ray index, so the order does not show.  (Regrouping per block of 256 rays
// instead needs four block barriers per round, and the warps with cheap classes wait at them for the warps with expensive
// ones: 45 % of the stall samples, profiles/r02_v6_sample_*.)
#define ODW_SAMPLE_CLASSES 8
#ifndef ODW_SAMPLE_SPAN
#define ODW_SAMPLE_SPAN 256     // measured on lambert-source, 2e7 rays binned in 2^21-ray waves: 64 / 128 / 256 / 512 / 1024 rays per warp 7.38 / 6.74 / 6.20 / 8.14 / 7.21 ms (512: fewer spans than resident warps; 1024: shared memory halves the occupancy)
#endif
static_assert(ODW_SAMPLE_CLASSES == 8 && ODW_SAMPLE_SPAN <= 8192, "class | rank << 3 must fit 16 bits");
__device__ __forceinline__ int emit_class(const DFace& f, int k) {
  return (f.flags & DFACE_TRI) ? 0 : 1 + (k % (ODW_SAMPLE_CLASSES - 1));
}

#ifndef ODW_SAMPLE_MINB
#define ODW_SAMPLE_MINB 4     // latency bound (table look-ups, fp64 trigonometry): 64 registers and 32 warps per SM measured 20 % faster than 128 and 16
#endif
__device__ __forceinline__ void sample_store(unsigned long long i, double first, double phi, const double* o, const double* d,
                                             double* first_out, double* phi_out, double* origins, double* dirs, unsigned int* keys, float bound) {
  if (first_out) first_out[i] = first;
  if (phi_out) phi_out[i] = phi;
  if (origins) { origins[3*i] = o[0]; origins[3*i+1] = o[1]; origins[3*i+2] = o[2]; }
  if (dirs) { dirs[3*i] = d[0]; dirs[3*i+1] = d[1]; dirs[3*i+2] = d[2]; }
  if (keys) keys[i] = ray_sort_key(o, d, bound);
}

__global__ void __launch_bounds__(256, ODW_SAMPLE_MINB) sample_kernel(DSource src, unsigned long long seed, unsigned long long first_ray,
                                                     unsigned long long n, double* first_out, double* phi_out,
                                                     double* origins, double* dirs, unsigned int* keys, float bound) {
  if (src.kind != ODW_SRC_SURFACE) {
    const unsigned long long stride = (unsigned long long)gridDim.x*blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {
      double u0, u1, first, phi, o[3], d[3];
      philox_uniform2(seed, (uint32_t)src.source_id, first_ray + i, 0u, u0, u1);
      sample_source(src, u0, u1, first, phi);
      make_ray(src, first, phi, o, d);
      sample_store(i, first, phi, o, d, first_out, phi_out, origins, dirs, keys, bound);
    }
    return;
  }
  // per warp: faces in ray order, faces in class order (4 B each), class | rank << 3 in ray order, ray of the span in class order (2 B each)
  extern __shared__ __align__(16) unsigned char s_sample[];
  __shared__ int s_count[8][ODW_SAMPLE_CLASSES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int* s_face_in = reinterpret_cast<unsigned int*>(s_sample) + (size_t)warp*2*ODW_SAMPLE_SPAN;
  unsigned int* s_face = s_face_in + ODW_SAMPLE_SPAN;
  unsigned short* s_key = reinterpret_cast<unsigned short*>(s_sample + (size_t)8*8*ODW_SAMPLE_SPAN) + (size_t)warp*2*ODW_SAMPLE_SPAN;
  unsigned short* s_ray = s_key + ODW_SAMPLE_SPAN;
  const unsigned long long stride = (unsigned long long)gridDim.x*(blockDim.x >> 5)*ODW_SAMPLE_SPAN;
  for (unsigned long long base = ((unsigned long long)blockIdx.x*(blockDim.x >> 5) + warp)*ODW_SAMPLE_SPAN; base < n; base += stride) {
    const int span = (int)min((unsigned long long)ODW_SAMPLE_SPAN, n - base);
    if (lane < ODW_SAMPLE_CLASSES) s_count[warp][lane] = 0;
    __syncwarp();
    for (int t = lane; t < span; t += 32) {
      double a0, a1;
      philox_uniform2(seed, (uint32_t)src.source_id, first_ray + base + t, 0u, a0, a1);
      const int face = pick_emit_face(src, a0);
      const int cls = emit_class(src.emit_faces[face], face);
      const int rank = atomicAdd(&s_count[warp][cls], 1);
      s_face_in[t] = (unsigned int)face; s_key[t] = (unsigned short)(cls | (rank << 3));
    }
    __syncwarp();
    if (lane == 0) { int acc = 0; for (int c = 0; c < ODW_SAMPLE_CLASSES; ++c) { const int m = s_count[warp][c]; s_count[warp][c] = acc; acc += m; } }
    __syncwarp();
    for (int t = lane; t < span; t += 32) {
      const int key = s_key[t], slot = s_count[warp][key & 7] + (key >> 3);
      s_face[slot] = s_face_in[t]; s_ray[slot] = (unsigned short)t;
    }
    __syncwarp();
    for (int t = lane; t < span; t += 32) {
      const unsigned long long i = base + s_ray[t];
      double first, phi;
      const RayInit r = init_ray_surface(src, seed, first_ray + i, &first, &phi, (int)s_face[t]);
      sample_store(i, first, phi, r.o, r.d, first_out, phi_out, origins, dirs, keys, bound);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// launch helpers used by odw_api.cu

// Kernel instances: the Monte-Carlo kernel for scenes staged in shared memory (the hot configuration) exists in four
// feature sets, everything else only with all features.  pick_feat() maps the features a launch needs to the leanest
// instance that covers them.
static int pick_feat(bool mc, bool bvh, int need) {
  if (bvh) return FEAT_ALL;
  if (!mc) return (need & FEAT_EXT) ? FEAT_ALL : FEAT_ALL & ~FEAT_EXT;   // explicit rays: fans, replays, pre-sampled surface sources
  if (need == 0) return 0;
  if ((need & ~FEAT_SEQ) == 0) return FEAT_SEQ;
  if (!(need & FEAT_EXT)) return FEAT_ALL & ~FEAT_EXT;      // surface sources / device binning in scenes of ideal surfaces (BASELINE configs[4])
  return FEAT_ALL;
}

template <bool MC, bool BVH, int FEAT>
static cudaError_t launch_instance(const TraceParams* p, int blocks, size_t smem, cudaStream_t st) {
  if (!BVH) cudaFuncSetAttribute(trace_kernel<MC, BVH, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  trace_kernel<MC, BVH, FEAT><<<blocks, ODW_THREADS, BVH ? 0 : smem, st>>>(*p);
  return cudaGetLastError();
}

template <bool MC, bool BVH, int FEAT>
static int occupancy_instance(size_t smem) {
  int nb = 0;
  if (!BVH) cudaFuncSetAttribute(trace_kernel<MC, BVH, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<MC, BVH, FEAT>, ODW_THREADS, BVH ? 0 : smem);
  return nb;
}

extern "C" cudaError_t odw_launch_trace(const TraceParams* p, bool mc, bool bvh, int need, int blocks, size_t smem, cudaStream_t st) {
  const int feat = pick_feat(mc, bvh, need);
  if (mc && !bvh) {
    if (feat == 0) return launch_instance<true, false, 0>(p, blocks, smem, st);
    if (feat == FEAT_SEQ) return launch_instance<true, false, FEAT_SEQ>(p, blocks, smem, st);
    if (feat == (FEAT_ALL & ~FEAT_EXT)) return launch_instance<true, false, FEAT_ALL & ~FEAT_EXT>(p, blocks, smem, st);
    return launch_instance<true, false, FEAT_ALL>(p, blocks, smem, st);
  }
  if (mc) return launch_instance<true, true, FEAT_ALL>(p, blocks, smem, st);
  if (bvh) return launch_instance<false, true, FEAT_ALL>(p, blocks, smem, st);
  if (feat == (FEAT_ALL & ~FEAT_EXT)) return launch_instance<false, false, FEAT_ALL & ~FEAT_EXT>(p, blocks, smem, st);
  return launch_instance<false, false, FEAT_ALL>(p, blocks, smem, st);
}

extern "C" cudaError_t odw_launch_sample(const DSource* src, unsigned long long seed, unsigned long long first_ray,
                                         unsigned long long n, double* first_out, double* phi_out, double* origins,
                                         double* dirs, unsigned int* keys, float bound, int blocks, cudaStream_t st) {
  const size_t smem = src->kind == ODW_SRC_SURFACE ? (size_t)8*12*ODW_SAMPLE_SPAN : 0;      // 8 warps x (2 x 4 B + 2 x 2 B) per ray of a span
  if (smem > 40*1024) cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sample_kernel<<<blocks, 256, smem, st>>>(*src, seed, first_ray, n, first_out, phi_out, origins, dirs, keys, bound);
  return cudaGetLastError();
}

extern "C" int odw_trace_threads(void) { return ODW_THREADS; }

extern "C" int odw_trace_occupancy(bool mc, bool bvh, int need, size_t smem) {
  const int feat = pick_feat(mc, bvh, need);
  if (mc && !bvh) {
    if (feat == 0) return occupancy_instance<true, false, 0>(smem);
    if (feat == FEAT_SEQ) return occupancy_instance<true, false, FEAT_SEQ>(smem);
    if (feat == (FEAT_ALL & ~FEAT_EXT)) return occupancy_instance<true, false, FEAT_ALL & ~FEAT_EXT>(smem);
    return occupancy_instance<true, false, FEAT_ALL>(smem);
  }
  if (mc) return occupancy_instance<true, true, FEAT_ALL>(smem);
  if (bvh) return occupancy_instance<false, true, FEAT_ALL>(smem);
  if (feat == (FEAT_ALL & ~FEAT_EXT)) return occupancy_instance<false, false, FEAT_ALL & ~FEAT_EXT>(smem);
  return occupancy_instance<false, false, FEAT_ALL>(smem);
}

extern "C" int odw_trace_instance(bool mc, bool bvh, int need) { return pick_feat(mc, bvh, need); }
