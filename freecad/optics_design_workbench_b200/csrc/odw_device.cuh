// odw_device.cuh — device-side geometry, optics and sampling for the sm_100a trace kernels.
//
// Semantics follow the reference (freecad/optics_design_workbench/freecad_elements/ray.py, cited per
// function); the arithmetic OCC does for the reference (line/surface intersection, point-on-trimmed-face,
// normal) is written out in closed form for plane / cylinder / cone / sphere / torus.  Everything is fp64:
// the acceptance rules work at distTol = 1e-6 mm on scenes of 1e2..1e3 mm.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../../include/odw.h"

#define ODW_TWO_PI 6.283185307179586476925286766559

// Face record as the kernels read it (built once per scene from odw_face).
struct DFace {
  double o[3], x[3], y[3], z[3];
  double p0, p1;
  double umin, umax, vmin, vmax;
  double aux[6];                   // conicoid with an ODW_SEG_ASPHERE record: aux[0..4] = coefficients of rho^4 .. rho^12, aux[5] = 1;
                                   // DFACE_FAST sphere zone / cylinder band: aux[0], aux[1] = axial share of the tolerance at c0, c1
  int32_t kind, trim, nsign, group;
  int32_t seg_first, seg_count, face_id, flags;
  unsigned long long seqmask[2];   // bit s set <=> the face's group is in SequentialModeElements step s
  double c0, c1, c2;               // fast-path constants: plane (o.z, o.x, o.y) | sphere/cylinder axial bounds (lo, hi)
  int32_t shell, pad_c;            // index of the face's shell (DScene::shells)
};
#define DFACE_FULL_U 1             // u range spans the whole period: no azimuth test needed
#define DFACE_FAST   2             // plane/uvbox, plane/disc, sphere (whole or full-u cap), cylinder (full-u band): inline test
#define DFACE_DISC   4             // plane whose only trim loop is one full circle: umin,vmin = centre, umax = radius
#define DFACE_PAIR   16            // a DFACE_FAST plane whose NEXT face (same shell) is a DFACE_FAST plane with the same normal: opposite
                                   // faces of a box share the reciprocal and the two dot products of the plane test (find_nearest_smem)
#define DFACE_TRI    8             // EMITTING faces only: plane trimmed to one triangle (a tessellated emitter), aux[0..5] = its corners (u, v)

// Shell record: the first-level cull of ray.py:345-374 (shell BoundBox enlarged by distTol).  The box is fp32,
// rounded outward; the kernel widens it by TraceParams::cull_margin (distTol + the fp32 error bound of the slab
// test) while staging it into shared memory, so the fp32 test can only over-accept, never reject a true hit.
struct DShell {                    // 64 B
  float lo[3], hi[3];
  int32_t face_first, face_count;
  unsigned long long seqmask[2];
  int32_t group;
  int32_t convex;                  // the shell bounds a convex solid: a ray leaving its surface outwards cannot hit it again
  int32_t pad[2];
};

struct DGroup {
  double n, reflectivity, absorption_length, lpm, order, gdir[3];
  int32_t type, record, gtype, fresnel;         // fresnel: opt-in Fresnel reflection at the faces of a Lens group (odw.h)
  int32_t scat_main, scat_modify, pad1, pad2;   // stochastic surface model: indices into DScene::scatters, -1 = ideal
};

// tabulated (theta, phi) density of a stochastic surface model; field names as in DSource so that the samplers are shared
struct DScatter {
  const double* phi_cdf;
  const double* first_cdf;
  const uint32_t* phi_guide;
  const uint32_t* first_guide;
  double first_lo, first_hi, phi_lo, phi_hi;
  int32_t n_first, n_phi, n_rows, n_tables;      // n_tables > 1: family over the incidence angle of the hit (odw.h odw_scatter)
};

struct DBinning {
  double origin[3], ua[3], va[3];
  double u_lo, v_lo, u_scale, v_scale, u_hi, v_hi;   // scale = n/(hi-lo)
  int32_t group, nu, nv, weighted;
  unsigned long long offset;                         // into the concatenated bins array
};

struct BvhNode {                   // builder node (host only)
  float lo[3], hi[3];
  int32_t left;                    // inner: index of left child (right = left+1); leaf: first primitive
  int32_t count;                   // 0 = inner node, >0 = number of primitives in the leaf
};

// Device node: an inner node carries the fp32 boxes of BOTH children (64 B = four 128-bit loads), so one fetch
// decides both children and their near/far order.  child: inner -> node index, leaf -> first entry of bvh_prims;
// count: 0 = inner, > 0 = leaf size, < 0 = no child.  Boxes are stored rounded outward and already widened by the
// launch's culling margin (odw_api.cu ensure_bvh_margin).
struct BvhNode2 {
  float lo0[3], hi0[3], lo1[3], hi1[3];
  int32_t child[2], count[2];
};

// 4-wide node of the wavefront traversal (odw_wavefront.cu): the fp32 boxes of up to four children, axis by axis, and their
// references.  128 B = eight 16-byte words: 0..5 = lo.x, hi.x, lo.y, hi.y, lo.z, hi.z of the four children, 6 = references,
// 7 unused.  An empty slot has lo = +3e38, hi = -3e38 (never hit).  Reference: >= 0 inner node; <= -2: leaf, -2 - ref = first entry
// of bvh4_prims << 3 | (entries - 1) (-1 = the traversal's "done").  A prims entry >= 0 is a face index (DScene::faces, general test); < 0 is ~index into
// the compact table of whole spheres (DScene::spheres).  Boxes are stored rounded outward and already widened by the launch's
// culling margin, like BvhNode2.
struct Bvh4Node {
  float b[6][4];                   // words 0..5: lo.x, hi.x, lo.y, hi.y, lo.z, hi.z (near / far plane of an axis = word 2a ^ (direction < 0))
  int32_t ref[4];
  int32_t pad[4];
};

// Whole sphere (surface kind sphere, no trim: the unit spheres of hugeArray's Draft arrays) as the traversal's compact leaf:
// 32 B instead of the 272 B face record; face index and optical group sit in DScene::sphere_info.
struct DSphere { double cx, cy, cz, r; };

struct DScene {
  const Bvh4Node* bvh4;            // wavefront traversal: 4-wide tree over the same face boxes (nullptr: none)
  const int32_t* bvh4_prims;
  const DSphere* spheres;
  const int2* sphere_info;         // (face index, optical group) of each compact sphere
  const ulonglong2* group_seqmask; // [n_groups]: bit s set <=> the group is in SequentialModeElements step s (the faces' seqmask, per group)
  int32_t n_bvh4_nodes, n_bvh4_prims, n_spheres, bvh_depth;
  const DFace* faces;
  const DShell* shells;
  const odw_trimseg* segs;
  const DGroup* groups;
  const DScatter* scatters;        // [n_scatters] stochastic surface models (nullptr: all surfaces ideal)
  const BvhNode2* bvh;             // nullptr: shells + faces staged in shared memory
  const int32_t* bvh_prims;        // face indices in leaf order
  int32_t n_faces, n_shells, n_segs, n_groups, n_seq_steps, n_bvh_nodes;
};

struct DSource {
  const double* phi_cdf;
  const double* first_cdf;
  const uint32_t* phi_guide;       // [GUIDE+1]
  const uint32_t* first_guide;     // [n_rows][GUIDE+1]
  double first_lo, first_hi, phi_lo, phi_hi, focal, wavelength;
  double M[12];                    // rows 0..2 of gpM
  int32_t kind, source_id, n_first, n_phi, n_rows, n_emit;
  // surface sources (ODW_SRC_SURFACE): emitting faces in the world frame, their trim loops, cumulative area weights
  const DFace* emit_faces;
  const odw_trimseg* emit_segs;
  const double* emit_cdf;
  const uint32_t* emit_guide;      // [ODW_EMIT_GUIDE+1]: emit_guide[k] = first face whose cumulative weight exceeds k/ODW_EMIT_GUIDE
  double dist_tol;
  int32_t n_first_guide, pad_;     // cells of first_guide of a surface source (a power of two near n_first: a 1e6-entry theta table is searched in <= 2 probes)
};
#define ODW_EMIT_GUIDE 65536
#define ODW_PRESAMPLE_SORT_BEGIN 8      // pre-sampled surface-source waves are traced in the order of these bits of ray_sort_key (24 bits = 3 radix passes)
#define ODW_PRESAMPLE_SORT_END 32
#define ODW_GUIDE 4096
#define ODW_BVH_STACK 64           // entries of a traversal stack (the builders bound the tree depth accordingly)

struct HitBuffers {
  double* points; double* dirs; double* powers;
  uint8_t* entering; unsigned long long* ray_index;
  int32_t* group; int32_t* bounce; int32_t* face_id; int32_t* medium;
  unsigned long long capacity;
};

struct Counters {                  // device-side odw_counts
  unsigned long long segments, hits, hits_dropped, escaped, depth_terminated, alive_next, dbg_cycles, dbg_ns;
};

struct TraceParams {
  DScene scene;
  DSource src;
  HitBuffers hits;
  Counters* counters;
  const DBinning* binnings; double* bins; int32_t n_binnings;
  // explicit ray input (nullptr for MC)
  const double* in_origins; const double* in_dirs; const double* in_powers;
  const unsigned int* in_order;      // explicit lists: the k-th claimed ray is row in_order[k] (coherence order of a pre-sampled wave), nullptr = row k
  // per-ray summary (explicit lists)
  int32_t* out_nseg; double* out_final_point; double* out_final_power; int32_t* out_final_medium;
  unsigned long long ignore_mask[4];
  unsigned long long seed, first_ray, n_rays;
  unsigned long long* ray_counter;   // next unclaimed ray of this launch (zeroed by the host): warps claim rays in chunks, see trace_kernel
  double max_len, tol, power_tol, wavelength;
  int32_t max_isect, sequential, record_all, store_hits;
  float cull_margin;                 // widening of the fp32 shell / BVH boxes, see odw_api.cu cull_margin()
  float origin_bound;                // max |coordinate| of any ray origin of this launch (host-side bound)
};

#ifdef ODW_DEVICE_CODE   // device functions: only the kernel translation units define this
namespace {              // internal linkage: two translation units (odw_kernels.cu, odw_wavefront.cu) include these definitions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
__device__ __forceinline__ double dot3(double ax, double ay, double az, const double* b) { return ax*b[0]+ay*b[1]+az*b[2]; }
__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) { return ax*bx+ay*by+az*bz; }

// Coherence key: rays that start in the same region and point the same way walk the same BVH nodes, so a warp of
// neighbours in key order stays converged.  12 bits origin cell (16^3 grid, Morton order) above 20 bits direction (octahedral map,
// 1024 x 1024, Morton order: a narrow beam from one point still spreads over thousands of direction cells).  Measured on hugeArray with the rays of the FIRST bounce sorted by direction on the host:
// 1.98e9 -> 3.11e9 segments/s (tools/gpu_coherence_probe.py).
#define ODW_SORT_KEY_BITS 32
__device__ __forceinline__ unsigned int spread2(unsigned int x) {     // 10 bits -> every second bit
  x = (x | (x << 8)) & 0x00ff00ffu; x = (x | (x << 4)) & 0x0f0f0f0fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__device__ __forceinline__ unsigned int spread3(unsigned int x) {     // 4 bits -> every third bit
  x = (x | (x << 4)) & 0x0c3u; x = (x | (x << 2)) & 0x249u;
  return x;
}
__device__ __forceinline__ unsigned int ray_sort_key(const double* point, const double* dn, float bound) {
  const float sc = 8.0f/bound;
  const int cx = min(15, max(0, (int)(((float)point[0] + bound)*sc)));
  const int cy = min(15, max(0, (int)(((float)point[1] + bound)*sc)));
  const int cz = min(15, max(0, (int)(((float)point[2] + bound)*sc)));
  const float dx = (float)dn[0], dy = (float)dn[1], dz = (float)dn[2];
  const float l1 = 1.0f/(fabsf(dx) + fabsf(dy) + fabsf(dz) + 1e-30f);
  float px = dx*l1, py = dy*l1;
  if (dz < 0) { const float qx = (1.0f - fabsf(py))*(px >= 0 ? 1.0f : -1.0f), qy = (1.0f - fabsf(px))*(py >= 0 ? 1.0f : -1.0f); px = qx; py = qy; }
  const unsigned int ux = (unsigned int)min(1023, max(0, (int)((px + 1.0f)*512.0f)));
  const unsigned int uy = (unsigned int)min(1023, max(0, (int)((py + 1.0f)*512.0f)));
  const unsigned int cell = spread3((unsigned int)cx) | (spread3((unsigned int)cy) << 1) | (spread3((unsigned int)cz) << 2);
  return (cell << 20) | spread2(ux) | (spread2(uy) << 1);
}

// Reciprocal / square root without the IEEE slow paths: MUFU seed (2^-23) + two Newton steps (~1 ulp).
// The correctly rounded '/' and sqrt() cost ~3x the instructions and, inlined at every face test, pushed the
// hot loop out of the instruction cache.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0); r = fma(r, e, r);
  e = fma(-x, r, 1.0); r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double h = 0.5*x;
  double e = fma(-h*r, r, 0.5); r = fma(r, e, r);
  e = fma(-h*r, r, 0.5); r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double fast_sqrt(double x) {          // x >= 0
  if (x <= 0) return 0.0;
  double r = fast_rsqrt(x);
  double sq = x*r;
  return fma(fma(-sq, sq, x), 0.5*r, sq);                          // one Heron correction on the product
}

// ---- Philox4x32-10, counter (ray_lo, ray_hi, source_id, purpose), key = seed -------------
__device__ __forceinline__ void philox_uniform2(unsigned long long seed, uint32_t source_id, unsigned long long ray,
                                                uint32_t purpose, double& u0, double& u1) {
  uint32_t c0 = (uint32_t)ray, c1 = (uint32_t)(ray >> 32), c2 = source_id, c3 = purpose;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u*c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u*c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  unsigned long long a = ((unsigned long long)c0 << 32) | c1, b = ((unsigned long long)c2 << 32) | c3;
  u0 = (double)(a >> 11)*(1.0/9007199254740992.0);
  u1 = (double)(b >> 11)*(1.0/9007199254740992.0);
}

// ---- sampler (reference distributions/random_number_generator.py:413-456,492-500) --------
__device__ __forceinline__ double linspace_at(double lo, double hi, int n, int i) {
  return (i >= n-1) ? hi : lo + (double)i*((hi-lo)/(double)(n-1));
}

// numpy.interp(x, cdf, linspace(lo,hi,n)); guide[k] = last index with cdf[j] <= k/G (G+1 entries; the theta table of a surface
// source, up to 1e6 entries, has G = DSource::n_first_guide)
__device__ __forceinline__ double interp_cdf(double x, const double* __restrict__ cdf, const uint32_t* __restrict__ guide,
                                             int n, double lo, double hi, const int G = ODW_GUIDE) {
  int k = (int)(x*(double)G);
  if (k > G-1) k = G-1;
  int a = (int)guide[k], b = (int)guide[k+1] + 1;     // cdf[a] <= x ; answer <= guide[k+1]
  if (b > n) b = n;
  while (b - a > 1) {
    int m = (a + b) >> 1;
    if (__ldg(cdf + m) <= x) a = m; else b = m;
  }
  if (a >= n-1) return hi;
  double xa = __ldg(cdf + a), xb = __ldg(cdf + a + 1);
  double fa = linspace_at(lo, hi, n, a);
  if (xa == x) return fa;
  double fb = linspace_at(lo, hi, n, a+1);
  double slope = (fb - fa)/(xb - xa);
  return slope*(x - xa) + fa;
}

__device__ __forceinline__ int nearest_row(double phi, double lo, double hi, int n_edges) {
  int nrows = n_edges - 1;
  double step = (hi - lo)/(double)(n_edges - 1);
  int k = (int)floor((phi - lo)/step);
  k = max(0, min(nrows-1, k));
  int best = -1; double bestd = 0;
  for (int i = k-2; i <= k+2; ++i) {
    if (i < 0 || i >= nrows) continue;
    double c = (linspace_at(lo, hi, n_edges, i+1) + linspace_at(lo, hi, n_edges, i))/2;
    double d = fabs(c - phi);
    if (best < 0 || d < bestd) { best = i; bestd = d; }
  }
  return best;
}

template <class Table>   // DSource or DScatter
__device__ __forceinline__ void sample_source(const Table& s, double u_phi, double u_first, double& first, double& phi) {
  phi = interp_cdf(u_phi, s.phi_cdf, s.phi_guide, s.n_phi, s.phi_lo, s.phi_hi);
  int row = 0;
  if (s.n_rows > 1) row = nearest_row(phi, s.phi_lo, s.phi_hi, s.n_phi);
  first = interp_cdf(u_first, s.first_cdf + (size_t)row*s.n_first, s.first_guide + (size_t)row*(ODW_GUIDE+1),
                     s.n_first, s.first_lo, s.first_hi);
}

// PointSourceProxy._makeRay (reference freecad_elements/point_source.py:411-460)
__device__ __forceinline__ void make_ray(const DSource& s, double first, double phi, double* o, double* d) {
  double lo[3], ld[3];
  double sp, cp; sincos(phi, &sp, &cp);
  if (s.kind == ODW_SRC_POINT_SPHERICAL) {
    double st, ct; sincos(first, &st, &ct);
    ld[0] = st*sp; ld[1] = -st*cp; ld[2] = ct;
    lo[0] = (0.0 - ld[0])*s.focal; lo[1] = (0.0 - ld[1])*s.focal; lo[2] = (1.0 - ld[2])*s.focal;
  } else {
    ld[0] = 0; ld[1] = 0; ld[2] = 1;
    lo[0] = first*cp; lo[1] = -first*sp; lo[2] = 0;
  }
  double l = sqrt(dot3(ld, ld));
  double q[3] = { lo[0]+ld[0]/l, lo[1]+ld[1]/l, lo[2]+ld[2]/l };
  double p1[3], p2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    p1[i] = s.M[4*i+0]*lo[0] + s.M[4*i+1]*lo[1] + s.M[4*i+2]*lo[2] + s.M[4*i+3];
    p2[i] = s.M[4*i+0]*q[0]  + s.M[4*i+1]*q[1]  + s.M[4*i+2]*q[2]  + s.M[4*i+3];
  }
  double dx = p2[0]-p1[0], dy = p2[1]-p1[1], dz = p2[2]-p1[2];
  double dl = sqrt(dx*dx + dy*dy + dz*dz);
  o[0] = p1[0]; o[1] = p1[1]; o[2] = p1[2];
  d[0] = dx/dl; d[1] = dy/dl; d[2] = dz/dl;
}

// ---- polynomial solvers -------------------------------------------------------------------
__device__ __forceinline__ int solve_quadratic(double a, double b, double c, double* t) {
  if (a == 0) { if (b == 0) return 0; t[0] = -c/b; return 1; }
  double disc = b*b - 4*a*c;
  if (disc < 0) return 0;
  double sq = sqrt(disc);
  double q = -0.5*(b + (b >= 0 ? sq : -sq));
  t[0] = q/a;
  t[1] = (q != 0) ? c/q : 0.0;
  return 2;
}

__device__ __noinline__ int solve_cubic_depressed(double p, double q, double* y) {
  int n = 0;
  double disc = q*q/4 + p*p*p/27;
  if (disc > 0) {
    double sq = sqrt(disc);
    y[n++] = cbrt(-q/2 + sq) + cbrt(-q/2 - sq);
  } else if (p == 0) {
    y[n++] = cbrt(-q);
  } else {
    double m = 2*sqrt(-p/3);
    double arg = 3*q/(p*m);
    arg = fmin(1.0, fmax(-1.0, arg));
    double th = acos(arg)/3;
    for (int k = 0; k < 3; ++k) y[n++] = m*cos(th - ODW_TWO_PI*k/3);
  }
  for (int i = 0; i < n; ++i)
    for (int it = 0; it < 3; ++it) {
      double f = (y[i]*y[i] + p)*y[i] + q, df = 3*y[i]*y[i] + p;
      if (df != 0) y[i] -= f/df;
    }
  return n;
}

// real roots of t^4 + B t^2 + C t + D in [lo, hi]; monotone pieces between the critical points
__device__ __noinline__ int solve_quartic_depressed(double B, double C, double D, double lo, double hi, double* roots) {
  double crit[3];
  int nc = solve_cubic_depressed(B/2, C/4, crit);
  for (int i = 0; i < nc; ++i) for (int j = i+1; j < nc; ++j)
    if (crit[j] < crit[i]) { double t = crit[i]; crit[i] = crit[j]; crit[j] = t; }
  double knots[5]; int nk = 0;
  knots[nk++] = lo;
  for (int i = 0; i < nc; ++i) if (crit[i] > lo && crit[i] < hi) knots[nk++] = crit[i];
  knots[nk++] = hi;
  int n = 0;
  for (int i = 0; i+1 < nk; ++i) {
    double a = knots[i], b = knots[i+1];
    double fa = ((a*a + B)*a + C)*a + D, fb = ((b*b + B)*b + C)*b + D;
    if (fa == 0) { roots[n++] = a; continue; }
    if (i+2 == nk && fb == 0) { roots[n++] = b; continue; }
    if ((fa > 0) == (fb > 0)) continue;
    double x = 0.5*(a+b);
    for (int it = 0; it < 200; ++it) {
      double f = ((x*x + B)*x + C)*x + D;
      if ((f > 0) == (fa > 0)) { a = x; fa = f; } else { b = x; fb = f; }
      double df = (4*x*x + 2*B)*x + C;
      double xn = (df != 0) ? x - f/df : 0.5*(a+b);
      if (!(xn > a && xn < b)) xn = 0.5*(a+b);
      if (fabs(xn - x) <= 1e-16*fmax(1.0, fabs(x))) { x = xn; break; }
      x = xn;
    }
    roots[n++] = x;
  }
  return n;
}

// ---- line / untrimmed surface (ray.py:411, infinite line, all points) ----------------------
__device__ __noinline__ int line_torus(const DFace& f, const double* w, const double* d, double* t) {
  double R = f.p0, r = f.p1;
  double t0 = -dot3(w, d);
  double o[3] = { w[0]+t0*d[0], w[1]+t0*d[1], w[2]+t0*d[2] };
  double m = dot3(o, o), rr = (R + r)*(R + r);
  if (m > rr) return 0;
  double half = sqrt(rr - m) + 1e-9;
  double oz = dot3(o, f.z), dz = dot3(d, f.z);
  double oxy2 = m - oz*oz, dxy2 = 1.0 - dz*dz, g = -oz*dz;
  double K = m + R*R - r*r;
  double roots[4];
  int n = solve_quartic_depressed(2*K - 4*R*R*dxy2, -8*R*R*g, K*K - 4*R*R*oxy2, -half, half, roots);
  for (int i = 0; i < n; ++i) t[i] = t0 + roots[i];
  return n;
}

// ---- even asphere on top of a conic of revolution (ODW_SEG_ASPHERE, include/odw.h); same arithmetic as the oracle ----
// sag as a function of u = rho^2:  S(u) = c u / (1 + sqrt(1 - (1+k) c^2 u)) + u^2 (a0 + a1 u + a2 u^2 + a3 u^3 + a4 u^4)
__device__ __forceinline__ bool asphere_sag(double c, double k, const double* a, double u, double& S, double& dSdu) {
  const double q2 = 1.0 - (1.0 + k)*c*c*u;
  if (!(q2 > 1e-14)) return false;                         // beyond (or on) the equator of the base conic
  const double q = sqrt(q2);
  S = c*u/(1.0 + q) + u*u*(a[0] + u*(a[1] + u*(a[2] + u*(a[3] + u*a[4]))));
  dSdu = c/(2.0*q) + u*(2*a[0] + u*(3*a[1] + u*(4*a[2] + u*(5*a[3] + u*6*a[4]))));
  return true;
}

// Newton on g(t) = z(t) - S(u(t)) with the ray reduced to the two coordinates the surface depends on:
// z(t) = wz + t dz (axial), u(t) = rho^2 = U0 + 2 U1 t + U2 t^2.  Few live values: this runs below the bounce loop in the
// call graph and every register it needs is one the loop cannot keep.  true = converged onto the surface.
__device__ __forceinline__ bool asphere_newton(const DFace& f, double U0, double U1, double U2, double wz, double dz, double t0, double& t_out) {
  double t = t0;
  for (int it = 0; it < 40; ++it) {
    double z = wz + t*dz, u = fmax(0.0, U0 + t*(2*U1 + U2*t)), S, dS;
    if (!asphere_sag(f.p0, f.p1, f.aux, u, S, dS)) return false;
    const double gp = dz - dS*2*(U1 + U2*t);
    if (gp == 0 || !isfinite(gp)) return false;
    const double dt = (z - S)/gp;
    t -= dt;
    if (fabs(dt) <= 1e-15*fmax(1.0, fabs(t))) {
      z = wz + t*dz; u = fmax(0.0, U0 + t*(2*U1 + U2*t));
      if (!asphere_sag(f.p0, f.p1, f.aux, u, S, dS)) return false;
      if (fabs(z - S) > 1e-10*(1.0 + fabs(z))) return false;
      t_out = t; return true;
    }
  }
  return false;
}

__device__ __noinline__ int line_surface(const DFace& f, const double* s, const double* d, double* t) {
  double w[3] = { s[0]-f.o[0], s[1]-f.o[1], s[2]-f.o[2] };
  switch (f.kind) {
    case ODW_SURF_PLANE: {
      double den = dot3(d, f.z);
      if (den == 0) return 0;
      t[0] = -dot3(w, f.z)/den;
      return 1;
    }
    case ODW_SURF_SPHERE:
      return solve_quadratic(1.0, 2*dot3(w, d), dot3(w, w) - f.p0*f.p0, t);
    case ODW_SURF_CYLINDER: {
      double wz = dot3(w, f.z), dz = dot3(d, f.z);
      double a = 1.0 - dz*dz;
      if (fabs(a) < 1e-300) return 0;
      return solve_quadratic(a, 2*(dot3(w, d) - wz*dz), dot3(w, w) - wz*wz - f.p0*f.p0, t);
    }
    case ODW_SURF_CONE: {
      double ta = tan(f.p1);
      double wz = dot3(w, f.z), dz = dot3(d, f.z);
      double r0 = f.p0 + wz*ta, r1 = dz*ta;
      return solve_quadratic(1.0 - dz*dz - r1*r1, 2*(dot3(w, d) - wz*dz - r0*r1), dot3(w, w) - wz*wz - r0*r0, t);
    }
    case ODW_SURF_TORUS:
      return line_torus(f, w, d, t);
    case ODW_SURF_CONICOID: {
      // c (rho^2 + (1+k) z^2) - 2 z = 0 along w + t d; a paraboloid met along its axis has a = 0 (one root).  Only the
      // sheet the sag formula describes counts: q = 1 - (1+k) c z >= 0.
      const double c = f.p0, k = f.p1;
      const double wz = dot3(w, f.z), dz = dot3(d, f.z);
      double r[2];
      const int n = solve_quadratic(c*(1.0 + k*dz*dz), 2*(c*(dot3(w, d) + k*wz*dz) - dz), c*(dot3(w, w) + k*wz*wz) - 2*wz, r);
      int m = 0;
      for (int i = 0; i < n; ++i) if (1.0 - (1.0 + k)*c*(wz + r[i]*dz) >= 0) t[m++] = r[i];
      return m;
    }
  }
  return 0;
}

// Conicoid with even-asphere terms: crossings of the base conic (or the vertex plane when it has none) refined by Newton.
// A SIBLING of line_surface in the call graph, not a callee: the register needs of the deepest call chain squeeze the
// bounce loop of the trace kernel (measured: as a callee of line_surface this cost the headline scene 17 %).
__device__ __noinline__ int line_asphere(const DFace& f, const double* s, const double* d, double* t) {
  double U0, U1, U2, wz, dz;
  {
    const double w[3] = { s[0]-f.o[0], s[1]-f.o[1], s[2]-f.o[2] };
    wz = dot3(w, f.z); dz = dot3(d, f.z);
    U2 = 1.0 - dz*dz; U1 = dot3(w, d) - wz*dz; U0 = dot3(w, w) - wz*wz;
  }
  const double c = f.p0, k1 = 1.0 + f.p1;
  double r[2];
  const int n = solve_quadratic(c*(U2 + k1*dz*dz), 2*(c*(U1 + k1*wz*dz) - dz), c*(U0 + k1*wz*wz) - 2*wz, r);
  int ns = 0;
  for (int i = 0; i < n; ++i) if (1.0 - k1*c*(wz + r[i]*dz) >= 0) r[ns++] = r[i];
  if (ns == 0 && dz != 0) { r[0] = -wz/dz; ns = 1; }
  int m = 0;
  for (int i = 0; i < ns; ++i) {
    double tt;
    if (!asphere_newton(f, U0, U1, U2, wz, dz, r[i], tt)) continue;
    if (m == 1 && fabs(tt - t[0]) <= 1e-9*fmax(1.0, fabs(tt))) continue;     // both starts found the same crossing
    t[m++] = tt;
  }
  return m;
}

// ---- point on trimmed face (ray.py:426, distToShape(face) < distTol) in (u,v) space ---------
__device__ __forceinline__ double wrap_into(double x, double lo) { return x - ODW_TWO_PI*floor((x - lo)/ODW_TWO_PI); }

__device__ __noinline__ bool loops_contain(const DFace& f, const odw_trimseg* __restrict__ segs, double u, double v,
                                           double su, double sv, double tol) {
  int crossings = 0;
  for (int i = 0; i < f.seg_count; ++i) {
    const odw_trimseg& s = segs[f.seg_first + i];
    const double* a = s.a;
    if (s.kind == ODW_SEG_LINE) {
      if ((a[1] > v) != (a[3] > v)) {
        double ux = a[0] + (v - a[1])*(a[2] - a[0])/(a[3] - a[1]);
        if (ux > u) ++crossings;
      }
      double ax = a[0]*su, ay = a[1]*sv, bx = a[2]*su, by = a[3]*sv, px = u*su, py = v*sv;
      double dx = bx-ax, dy = by-ay, l2 = dx*dx + dy*dy;
      double tt = l2 > 0 ? ((px-ax)*dx + (py-ay)*dy)/l2 : 0;
      tt = fmin(1.0, fmax(0.0, tt));
      double qx = ax + tt*dx - px, qy = ay + tt*dy - py;
      if (sqrt(qx*qx + qy*qy) < tol) return true;
    } else {
      double cu = a[0], cv = a[1], r = a[2], a0 = a[3], span = a[4];
      double dv = v - cv, du = u - cu;
      bool full = span >= ODW_TWO_PI - 1e-12;
      if (fabs(dv) < r) {
        double h = sqrt(r*r - dv*dv);
        if (full) {
          if (cu - h > u) ++crossings;
          if (cu + h > u) ++crossings;
        } else {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            double ux = k ? cu + h : cu - h;
            if (ux > u) {
              double rel = atan2(dv, ux - cu) - a0; rel -= ODW_TWO_PI*floor(rel/ODW_TWO_PI);
              if (rel <= span) ++crossings;
            }
          }
        }
      }
      double rad = sqrt(du*du + dv*dv);
      if (fabs(rad - r)*fmin(su, sv) < tol) {
        if (full) return true;
        double rel = atan2(dv, du) - a0; rel -= ODW_TWO_PI*floor(rel/ODW_TWO_PI);
        if (rel <= span) return true;
      }
    }
  }
  return (crossings & 1) != 0;
}

// P on the untrimmed surface of f.  returns true iff P lies on the trimmed face dilated by tol.
__device__ __noinline__ bool on_trimmed_face(const DFace& f, const odw_trimseg* __restrict__ segs, const double* P, double tol) {
  if (f.trim == ODW_TRIM_NONE) return true;
  double w[3] = { P[0]-f.o[0], P[1]-f.o[1], P[2]-f.o[2] };
  double x = dot3(w, f.x), y = dot3(w, f.y);
  double u, v, su = 1, sv = 1;
  bool uper = true, vper = false;
  switch (f.kind) {
    case ODW_SURF_PLANE: u = x; v = y; uper = false; break;
    case ODW_SURF_CYLINDER: v = dot3(w, f.z); su = f.p0; u = 0; break;
    case ODW_SURF_CONE: {
      double sa, ca; sincos(f.p1, &sa, &ca);
      v = dot3(w, f.z)/ca;
      double rs = f.p0 + v*sa;
      if (rs < 0) { x = -x; y = -y; }
      su = fabs(rs); u = 0; break;
    }
    case ODW_SURF_SPHERE: {
      double sn = fmin(1.0, fmax(-1.0, dot3(w, f.z)/f.p0));
      v = asin(sn); su = f.p0*sqrt(fmax(0.0, 1 - sn*sn)); sv = f.p0; u = 0; break;
    }
    case ODW_SURF_CONICOID: {   // v = rho; meridian arc length per unit rho = sqrt(1 + z'^2), z' = c rho / q
      v = sqrt(x*x + y*y);
      const double q2 = fmax(1e-12, 1.0 - (1.0 + f.p1)*f.p0*f.p0*v*v);
      double slope = f.p0*v/sqrt(q2);
      if (f.aux[5] != 0.0) { const double u2 = v*v; slope += 2*v*u2*(2*f.aux[0] + u2*(3*f.aux[1] + u2*(4*f.aux[2] + u2*(5*f.aux[3] + u2*6*f.aux[4])))); }
      su = v; sv = sqrt(1.0 + slope*slope); u = 0; break;
    }
    default: {   // torus
      double rho = sqrt(x*x + y*y);
      v = atan2(dot3(w, f.z), rho - f.p0);
      su = rho; sv = f.p1; vper = true; u = 0; break;
    }
  }
  su = fmax(su, 1e-12);
  double tu = tol/su, tv = tol/sv;
  if (vper) v = wrap_into(v, f.vmin - tv);
  if (v < f.vmin - tv || v > f.vmax + tv) return false;
  if (uper) {
    if (!(f.flags & DFACE_FULL_U) || f.trim == ODW_TRIM_LOOPS) {
      u = wrap_into(atan2(y, x), f.umin - tu);
      if (u < f.umin - tu || u > f.umax + tu) return false;
    }
  } else {
    if (u < f.umin - tu || u > f.umax + tu) return false;
  }
  if (f.trim == ODW_TRIM_UVBOX) return true;
  return loops_contain(f, segs, u, v, su, sv, tol);
}

// outward unit normal at P (ray.py:463-465 + face orientation)
__device__ __forceinline__ void outward_normal_general(const DFace& f, const double* P, double* n) {
  double w[3] = { P[0]-f.o[0], P[1]-f.o[1], P[2]-f.o[2] };
  double g[3];
  switch (f.kind) {
    case ODW_SURF_PLANE: g[0] = f.z[0]; g[1] = f.z[1]; g[2] = f.z[2]; break;
    case ODW_SURF_SPHERE: g[0] = w[0]; g[1] = w[1]; g[2] = w[2]; break;
    case ODW_SURF_CYLINDER: {
      double z = dot3(w, f.z);
      g[0] = w[0]-z*f.z[0]; g[1] = w[1]-z*f.z[1]; g[2] = w[2]-z*f.z[2]; break;
    }
    case ODW_SURF_CONE: {
      double sa, ca; sincos(f.p1, &sa, &ca);
      double z = dot3(w, f.z);
      double rx = w[0]-z*f.z[0], ry = w[1]-z*f.z[1], rz = w[2]-z*f.z[2];
      double rho = sqrt(rx*rx + ry*ry + rz*rz);
      double sg = (f.p0 + z/ca*sa) >= 0 ? 1.0 : -1.0;
      g[0] = ca*rx/rho - sg*sa*f.z[0]; g[1] = ca*ry/rho - sg*sa*f.z[1]; g[2] = ca*rz/rho - sg*sa*f.z[2]; break;
    }
    case ODW_SURF_CONICOID: {   // du x dv ~ c rho_vec - q Z, q = 1 - (1+k) c z
      const double z = dot3(w, f.z);
      const double rx = w[0]-z*f.z[0], ry = w[1]-z*f.z[1], rz = w[2]-z*f.z[2];
      double m = f.p0, zc = z;
      if (f.aux[5] != 0.0) {
        // even-asphere terms P(u), u = rho^2: the point minus P lies on the base conic at the same rho, which gives q without a
        // square root; slope = c rho / q + dP/drho  =>  direction (c + 2 q dP/du) rho_vec - q Z
        const double u = rx*rx + ry*ry + rz*rz;
        zc = z - u*u*(f.aux[0] + u*(f.aux[1] + u*(f.aux[2] + u*(f.aux[3] + u*f.aux[4]))));
        m = u*(2*f.aux[0] + u*(3*f.aux[1] + u*(4*f.aux[2] + u*(5*f.aux[3] + u*6*f.aux[4]))));
      }
      const double q = 1.0 - (1.0 + f.p1)*f.p0*zc;
      if (f.aux[5] != 0.0) m = f.p0 + 2*q*m;
      g[0] = m*rx - q*f.z[0]; g[1] = m*ry - q*f.z[1]; g[2] = m*rz - q*f.z[2]; break;
    }
    default: {
      double z = dot3(w, f.z);
      double rx = w[0]-z*f.z[0], ry = w[1]-z*f.z[1], rz = w[2]-z*f.z[2];
      double k = f.p0/sqrt(rx*rx + ry*ry + rz*rz);
      g[0] = w[0]-k*rx; g[1] = w[1]-k*ry; g[2] = w[2]-k*rz; break;
    }
  }
  double s = (double)f.nsign/sqrt(dot3(g, g));
  n[0] = g[0]*s; n[1] = g[1]*s; n[2] = g[2]*s;
}

// ---- surface source (reference freecad_elements/surface_source.py:85-111,390-410,522-555) ----------------
// point and first derivatives of the parametrisation (OCC's, see odw_face in include/odw.h)
__device__ __forceinline__ void surface_eval(const DFace& f, double u, double v, double* P, double* du, double* dv) {
  if (f.kind == ODW_SURF_PLANE) {
    for (int i = 0; i < 3; ++i) { P[i] = f.o[i] + u*f.x[i] + v*f.y[i]; du[i] = f.x[i]; dv[i] = f.y[i]; }
    return;
  }
  double su, cu; sincos(u, &su, &cu);
  double rad[3], tang[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { rad[i] = cu*f.x[i] + su*f.y[i]; tang[i] = -su*f.x[i] + cu*f.y[i]; }
  switch (f.kind) {
    case ODW_SURF_CYLINDER:
      for (int i = 0; i < 3; ++i) { P[i] = f.o[i] + f.p0*rad[i] + v*f.z[i]; du[i] = f.p0*tang[i]; dv[i] = f.z[i]; }
      break;
    case ODW_SURF_CONE: {
      double sa, ca; sincos(f.p1, &sa, &ca);
      const double r = f.p0 + v*sa;
      for (int i = 0; i < 3; ++i) { P[i] = f.o[i] + r*rad[i] + v*ca*f.z[i]; du[i] = r*tang[i]; dv[i] = sa*rad[i] + ca*f.z[i]; }
      break;
    }
    case ODW_SURF_SPHERE: {
      double sv, cv; sincos(v, &sv, &cv);
      const double R = f.p0;
      for (int i = 0; i < 3; ++i) { P[i] = f.o[i] + R*cv*rad[i] + R*sv*f.z[i]; du[i] = R*cv*tang[i]; dv[i] = -R*sv*rad[i] + R*cv*f.z[i]; }
      break;
    }
    default: {
      double sv, cv; sincos(v, &sv, &cv);
      const double R = f.p0, r = f.p1;
      for (int i = 0; i < 3; ++i) { P[i] = f.o[i] + (R + r*cv)*rad[i] + r*sv*f.z[i]; du[i] = (R + r*cv)*tang[i]; dv[i] = -r*sv*rad[i] + r*cv*f.z[i]; }
    }
  }
}

// (u, v) distributed by area inside the face's parameter window; false = draw rejected (cone, torus)
__device__ __forceinline__ bool surface_draw_uv(const DFace& f, double w0, double w1, double w2, double& u, double& v) {
  double u0 = f.umin, u1 = f.umax, v0 = f.vmin, v1 = f.vmax;
  if (f.trim == ODW_TRIM_NONE) {
    u0 = 0; u1 = ODW_TWO_PI;
    if (f.kind == ODW_SURF_SPHERE) { v0 = -ODW_TWO_PI/4; v1 = ODW_TWO_PI/4; }
    if (f.kind == ODW_SURF_TORUS) { v0 = 0; v1 = ODW_TWO_PI; }
  }
  u = u0 + w0*(u1 - u0);
  switch (f.kind) {
    case ODW_SURF_SPHERE: {
      const double s0 = sin(v0), s1 = sin(v1);
      v = asin(fmin(1.0, fmax(-1.0, s0 + w1*(s1 - s0))));
      return true;
    }
    case ODW_SURF_CONE: {
      const double sa = sin(f.p1), r0 = fabs(f.p0 + v0*sa), r1 = fabs(f.p0 + v1*sa);
      v = v0 + w1*(v1 - v0);
      return w2*fmax(r0, r1) < fabs(f.p0 + v*sa);
    }
    case ODW_SURF_TORUS:
      v = v0 + w1*(v1 - v0);
      return w2*(f.p0 + f.p1) < f.p0 + f.p1*cos(v);
    default:
      v = v0 + w1*(v1 - v0);
      return true;
  }
}

// ---- stochastic surface model (optical_group.py:279-323) -------------------------------------
// Rotation(n, phi) * Rotation(n x d_in, theta) * n = cos(theta) n^ + sin(theta) (cos(phi) (a x n^) + sin(phi) a), a = unit(n x d_in),
// scaled to |n|; a vanishing axis rotates nothing
__device__ __forceinline__ void scatter_direction(const double* n, const double* d_in, double theta, double phi, double* out) {
  const double nl = sqrt(dot3(n, n));
  const double nh[3] = { n[0]/nl, n[1]/nl, n[2]/nl };
  double a[3] = { nh[1]*d_in[2]-nh[2]*d_in[1], nh[2]*d_in[0]-nh[0]*d_in[2], nh[0]*d_in[1]-nh[1]*d_in[0] };
  const double al = sqrt(dot3(a, a));
  if (!(al > 1e-300)) { out[0] = n[0]; out[1] = n[1]; out[2] = n[2]; return; }
  a[0] /= al; a[1] /= al; a[2] /= al;
  const double axn[3] = { a[1]*nh[2]-a[2]*nh[1], a[2]*nh[0]-a[0]*nh[2], a[0]*nh[1]-a[1]*nh[0] };
  double st, ct, sp, cp; sincos(theta, &st, &ct); sincos(phi, &sp, &cp);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = nl*(ct*nh[i] + st*(cp*axn[i] + sp*a[i]));
}

// ---- Ray.mirror / snellsLaw / lineGrating (ray.py:482-539) ---------------------------------
__device__ __forceinline__ void mirror_dir(const double* ray, const double* n, double* out) {
  double k = 2*dot3(ray, n);
  out[0] = ray[0]-k*n[0]; out[1] = ray[1]-k*n[1]; out[2] = ray[2]-k*n[2];
}

__device__ __forceinline__ bool snell(const double* ray, double n1, double n2, const double* n, double* out) {
  double cx = n[1]*ray[2]-n[2]*ray[1], cy = n[2]*ray[0]-n[0]*ray[2], cz = n[0]*ray[1]-n[1]*ray[0];
  double mu = n1/n2;
  double root = 1 - mu*mu*(cx*cx + cy*cy + cz*cz);
  if (root < 0) { mirror_dir(ray, n, out); return true; }
  // n x ((-n) x ray) = ray (n.n) - n (n.ray)
  double nn = dot3(n, n), nr = dot3(n, ray), s = sqrt(root);
  out[0] = mu*(ray[0]*nn - n[0]*nr) + n[0]*s;
  out[1] = mu*(ray[1]*nn - n[1]*nr) + n[1]*s;
  out[2] = mu*(ray[2]*nn - n[2]*nr) + n[2]*s;
  return false;
}

// unpolarised Fresnel reflectance of the interface n1 -> n2 (unit direction, unit normal with d.n >= 0); 1 beyond the critical
// angle.  Opt-in extension (odw_group.fresnel): the reference has no Fresnel split.  Same arithmetic as the oracle.
__device__ __forceinline__ double fresnel_reflectance(const double* d, const double* n, double n1, double n2) {
  const double ci = fmin(1.0, fabs(dot3(d, n)));
  const double s2 = (n1/n2)*(n1/n2)*(1 - ci*ci);
  if (s2 >= 1) return 1.0;
  const double ct = sqrt(1 - s2);
  const double rs = (n1*ci - n2*ct)/(n1*ci + n2*ct), rp = (n1*ct - n2*ci)/(n1*ct + n2*ci);
  return 0.5*(rs*rs + rp*rp);
}

struct Vec3 { double x, y, z; };
// by value in, by value out: pointers to the caller's ray state would pin it to local memory
__device__ __noinline__ Vec3 line_grating(double rx, double ry, double rz, double n1, double n2, double nx, double ny, double nz,
                                          const DGroup* gp, double wavelength_nm, bool transmission) {
  const DGroup& g = *gp;
  const double ray_in[3] = { rx, ry, rz }, normal[3] = { nx, ny, nz };
  double out[3];
  double wl = wavelength_nm/1000.0;
  double rl = sqrt(dot3(ray_in, ray_in)), nl = sqrt(dot3(normal, normal)), gl = sqrt(dot3(g.gdir, g.gdir));
  double ray[3], sn[3], gv[3];
  for (int i = 0; i < 3; ++i) { ray[i] = ray_in[i]/rl; sn[i] = normal[i]/nl; gv[i] = g.gdir[i]/gl; }
  double P[3] = { gv[1]*sn[2]-gv[2]*sn[1], gv[2]*sn[0]-gv[0]*sn[2], gv[0]*sn[1]-gv[1]*sn[0] };
  double pl = sqrt(dot3(P, P)); P[0] /= pl; P[1] /= pl; P[2] /= pl;
  double D[3] = { sn[1]*P[2]-sn[2]*P[1], sn[2]*P[0]-sn[0]*P[2], sn[0]*P[1]-sn[1]*P[0] };
  double dl = sqrt(dot3(D, D)); D[0] /= dl; D[1] /= dl; D[2] /= dl;
  double mu = n1/n2, d = 1000.0/g.lpm;
  double T = (g.order*wl)/(n1*d);
  double nn = dot3(sn, sn);
  double V = (mu*dot3(ray, sn))/nn;
  double W = (mu*mu - 1 + T*T - 2*mu*T*dot3(ray, D))/nn;
  double sq = sqrt((2*V)*(2*V) - 4*W);
  double q1 = (-2*V + sq)/2, q2 = (-2*V - sq)/2;
  double Q = transmission ? fmin(q1, q2) : fmax(q1, q2);
  for (int i = 0; i < 3; ++i) out[i] = -(mu*ray[i] - T*D[i] + Q*sn[i]);
  Vec3 r; r.x = out[0]; r.y = out[1]; r.z = out[2];
  return r;
}
}  // namespace
#endif  // ODW_DEVICE_CODE
