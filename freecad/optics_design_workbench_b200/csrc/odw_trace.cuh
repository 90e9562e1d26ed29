// odw_trace.cuh — device code shared by the trace kernels: the exact face tests and the nearest-hit rule of
// Ray.findNearestIntersection (reference freecad_elements/ray.py:290-452), hit recording, ray generation and the
// surface interaction of Ray.traceRay (ray.py:105-281).  Included by odw_kernels.cu (register-resident kernel for
// scenes staged in shared memory) and odw_wavefront.cu (wavefront kernels for BVH scenes).
#pragma once
#include <cooperative_groups.h>
#define ODW_DEVICE_CODE
#include "odw_device.cuh"

namespace {   // internal linkage, see odw_device.cuh

namespace cg = cooperative_groups;

// Scene / launch features a kernel instance is compiled for.  The trace kernel is one big state machine; code for features
// the scene does not use still costs registers (spills) and instruction-cache room in the bounce loop, so the host picks
// the leanest instance that covers the launch (odw_api.cu launch_features): measured +12 % on lensesAndMirrors.
enum { FEAT_EXT = 1,        // gratings, stochastic surface models, finite absorption lengths, even-asphere faces
       FEAT_SURFSRC = 2,    // surface light source (init_ray_surface)
       FEAT_SEQ = 4,        // SequentialMode filtering
       FEAT_BIN = 8,        // detector binning on the device
       FEAT_ALL = 15 };

// per-CTA event counters in shared memory (flushed to the global Counters once per CTA): keeping them in registers
// cost five registers per lane for values touched once per ray
enum { CNT_SEGMENTS = 0, CNT_HITS, CNT_DROPPED, CNT_ESCAPED, CNT_DEPTH, CNT_N };

// outward unit normal: short paths for plane / sphere / cylinder, general otherwise
__device__ __forceinline__ void outward_normal(const DFace& f, const double* P, double* n) {
  if (f.kind == ODW_SURF_PLANE) {
    const double sg = (double)f.nsign;
    n[0] = sg*f.z[0]; n[1] = sg*f.z[1]; n[2] = sg*f.z[2];
  } else if (f.kind == ODW_SURF_SPHERE || f.kind == ODW_SURF_CYLINDER) {
    double g0 = P[0]-f.o[0], g1 = P[1]-f.o[1], g2 = P[2]-f.o[2];
    if (f.kind == ODW_SURF_CYLINDER) {
      const double z = dot3(g0, g1, g2, f.z);
      g0 -= z*f.z[0]; g1 -= z*f.z[1]; g2 -= z*f.z[2];
    }
    const double sc = (double)f.nsign*fast_rsqrt(g0*g0 + g1*g1 + g2*g2);
    n[0] = g0*sc; n[1] = g1*sc; n[2] = g2*sc;
  } else {
    outward_normal_general(f, P, n);
  }
}

struct NearestHit {
  double tA, tB;     // closest accepted hit overall / closest whose group differs from the current medium
  double lim;        // min(maxRayLength + distTol, tA + 2 distTol): nothing at or beyond it can be chosen (ray.py:425,432,440)
  int fA, fB;
};

// general face test: any surface kind, any trim (cone, torus, partial azimuth ranges, pcurve loops).  Out of line and
// called BY VALUE: a pointer to the caller's ray state would force that state into local memory for the whole
// kernel.  Returns the smallest t in (tol, lim) whose point lies on the trimmed face, or +inf.  (All hits of one face
// share its group, so only the nearest one can win either slot of NearestHit.)
template <int FEAT>
__device__ __noinline__ double general_nearest(const DFace* fp, const odw_trimseg* __restrict__ segs, double tol,
                                               double sx, double sy, double sz, double dx, double dy, double dz, double lim) {
  const DFace& f = *fp;
  const double s[3] = { sx, sy, sz }, dn[3] = { dx, dy, dz };
  if (f.kind == ODW_SURF_TORUS) {
    // Cheap exact rejects before the quartic: the torus lies between the planes |z| <= r and between the cylinders
    // R - r <= rho <= R + r of its own frame.  Where the ray is inside the slab (a t interval), rho^2(t) is a convex
    // parabola: its maximum over the interval is at an end, its minimum at the clamped vertex.  (A ray through the hole of
    // the ring — the last segment of every ray of lensesAndMirrors — is rejected here without solving anything.)
    const double w[3] = { s[0]-f.o[0], s[1]-f.o[1], s[2]-f.o[2] };
    const double wz = dot3(w, f.z), dz = dot3(dn, f.z), rr = f.p1 + tol;
    double ta = 0.0, tb = lim;
    if (fabs(dz) > 1e-300) {
      const double t1 = (-rr - wz)/dz, t2 = (rr - wz)/dz;
      ta = fmax(ta, fmin(t1, t2)); tb = fmin(tb, fmax(t1, t2));
    } else if (fabs(wz) > rr) return 1e300;
    if (!(ta <= tb)) return 1e300;
    const double A = 1.0 - dz*dz, B = dot3(w, dn) - wz*dz, C = dot3(w, w) - wz*wz;
    const double ra = (A*ta + 2*B)*ta + C, rb = (A*tb + 2*B)*tb + C;
    const double inner = f.p0 - rr, outer = f.p0 + rr;
    if (inner > 0 && fmax(ra, rb) < inner*inner) return 1e300;
    double tv = A > 1e-300 ? fmin(tb, fmax(ta, -B/A)) : ta;
    if ((A*tv + 2*B)*tv + C > outer*outer) return 1e300;
  }
  double ts[4];
  // even-asphere faces only in instances with FEAT_EXT: the Newton solver below this call would squeeze the registers of the
  // bounce loop of every instance that can reach it (measured: -17 % on the headline scene, which has no such face)
  int nt = ((FEAT & FEAT_EXT) && f.kind == ODW_SURF_CONICOID && f.aux[5] != 0.0) ? line_asphere(f, s, dn, ts) : line_surface(f, s, dn, ts);
  double best = 1e300;
  for (int k = 0; k < nt; ++k) {
    double t = ts[k];
    if (!(t > tol)) continue;                                        // ray.py:424  |P - start| > distTol (and forward)
    if (!(t < lim) || !(t < best)) continue;                         // ray.py:425,432,440
    double P[3] = { s[0]+t*dn[0], s[1]+t*dn[1], s[2]+t*dn[2] };
    if (!on_trimmed_face(f, segs, P, tol)) continue;                 // ray.py:426
    best = t;
  }
  return best;
}

__device__ __forceinline__ void accept_hit(double t, int idx, int group, int medium, double tol, NearestHit& h) {
  if (t < h.tA) { h.tA = t; h.fA = idx; h.lim = fmin(h.lim, t + 2*tol); }
  if (group != medium && t < h.tB) { h.tB = t; h.fB = idx; }
}

// trimmed part of the plane fast paths: the crossing at parameter t against the rectangle / disc dilated by the tolerance
__device__ __forceinline__ void plane_accept(const DFace& f, int idx, double t, const double* s, const double* dn, int medium, double tol, NearestHit& h) {
  if (t > tol && t < h.lim) {
    const double Px = fma(t, dn[0], s[0]), Py = fma(t, dn[1], s[1]), Pz = fma(t, dn[2], s[2]);
    const double u = dot3(Px, Py, Pz, f.x) - f.c1, v = dot3(Px, Py, Pz, f.y) - f.c2;
    if (f.flags & DFACE_DISC) {
      // one full circle as the only trim loop (every lens flat): distance of P to the disc < tol  <=>  rho < r + tol
      const double du = u - f.umin, dv = v - f.vmin, rr = f.umax + tol;
      if (du*du + dv*dv < rr*rr) accept_hit(t, idx, f.group, medium, tol, h);
    } else if (u >= f.umin - tol && u <= f.umax + tol && v >= f.vmin - tol && v <= f.vmax + tol) accept_hit(t, idx, f.group, medium, tol, h);
  }
}

// two planes with the same normal (opposite faces of a box, DFACE_PAIR): one reciprocal and one pair of dot products serve
// both; each t is bit-identical to what test_face computes for the face alone
__device__ __forceinline__ void test_plane_pair(const DFace& fa, const DFace& fb, int idx, const double* s, const double* dn, int medium, double tol, NearestHit& h) {
  const double r = fast_rcp(dot3(dn, fa.z)), sz = dot3(s, fa.z);
  const double ta = (fa.c0 - sz)*r, tb = (fb.c0 - sz)*r;
  plane_accept(fa, idx, ta, s, dn, medium, tol, h);
  plane_accept(fb, idx + 1, tb, s, dn, medium, tol, h);
}

// One face against the line start + t*dn (ray.py:407-432).  tmax = maxRayLength + distTol.
// Fast paths (inline, no division / inverse trigonometry): rectangle on a plane, whole sphere or spherical cap/zone
// with full azimuth, cylinder band with full azimuth.  Everything else goes through test_face_general.
template <bool CHECK_GROUP, int FEAT>
__device__ __forceinline__ void test_face(const DFace& f, int idx, const TraceParams& p, const double* s, const double* dn,
                                          int medium, int seq_index, double tmax, NearestHit& h) {
  if (CHECK_GROUP) {   // the shared-memory path filters whole shells instead
    if (p.sequential && (seq_index >= 128 || !((f.seqmask[seq_index >> 6] >> (seq_index & 63)) & 1ull))) return;
    if (f.group < 256 && ((p.ignore_mask[f.group >> 6] >> (f.group & 63)) & 1ull)) return;   // IgnoredOpticalElements
  }
  const double tol = p.tol;
  if (!(f.flags & DFACE_FAST)) {
    const double t = general_nearest<FEAT>(&f, p.scene.segs, tol, s[0], s[1], s[2], dn[0], dn[1], dn[2], h.lim);
    if (t < 1e299) accept_hit(t, idx, f.group, medium, tol, h);
    return;
  }
  const double limit = h.lim;
  if (f.kind == ODW_SURF_PLANE) {
    const double den = dot3(dn, f.z);
    const double t = (f.c0 - dot3(s, f.z))*fast_rcp(den);            // den == 0: inf/NaN fails the range test
    plane_accept(f, idx, t, s, dn, medium, tol, h);
    return;
  }
  // sphere / cylinder: a t^2 + 2 b t + c = 0 in the coordinates of the axis frame
  const double w0 = s[0]-f.o[0], w1 = s[1]-f.o[1], w2 = s[2]-f.o[2];
  double a, b, c;
  const double wz = dot3(w0, w1, w2, f.z), dz = dot3(dn, f.z);
  if (f.kind == ODW_SURF_SPHERE) {
    a = 1.0; b = dot3(w0, w1, w2, dn); c = dot3(w0, w1, w2, w0, w1, w2) - f.p0*f.p0;
  } else {
    a = 1.0 - dz*dz; b = dot3(w0, w1, w2, dn) - wz*dz; c = dot3(w0, w1, w2, w0, w1, w2) - wz*wz - f.p0*f.p0;
  }
  const double disc = b*b - a*c;
  if (!(disc >= 0) || a < 1e-300) return;
  const double sq = fast_sqrt(disc);
  const double q = -(b + (b >= 0 ? sq : -sq));                       // stable: no cancellation in q
  const double r0 = q*fast_rcp(a), r1 = (q != 0) ? c*fast_rcp(q) : 0.0;
  const double tn = fmin(r0, r1), tf = fmax(r0, r1);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double t = k ? tf : tn;
    if (t > tol && t < limit) {
      const double zc = fma(t, dz, wz);                              // axial coordinate of the hit
      if (zc >= fma(-tol, f.aux[0], f.c0) && zc <= fma(tol, f.aux[1], f.c1)) accept_hit(t, idx, f.group, medium, tol, h);   // aux: axial share of the tolerance
    }
  }
}

// Ray.findNearestIntersection (ray.py:290-452) on the scene staged in shared memory: shells are culled by
// their box first (ray.py:345-374), faces of surviving shells are tested one by one.  All lanes of a warp walk
// the same shell/face lists, so shared-memory reads are broadcasts.
// The shell cull is a conservative fp32 slab test (boxes widened by cull_margin while staging, see odw_api.cu):
// it only decides which faces get the exact fp64 test, so it cannot change a result.  In fp64 this cull was 45 %
// of all executed instructions (fmin/fmax on doubles are multi-instruction sequences; FMNMX is one).
template <int FEAT>
__device__ __forceinline__ int find_nearest_smem(const DShell* sshells, const DFace* sfaces, const TraceParams& p,
                                                 const double* s, const double* dn,
                                                 int medium, int seq_index, int skip_shell, double max_len, double& t_out) {
  const double tol = p.tol;
  const double tmax = max_len + tol;
  NearestHit h; h.tA = 1e300; h.tB = 1e300; h.lim = tmax; h.fA = -1; h.fB = -1;
  const float sx = (float)s[0], sy = (float)s[1], sz = (float)s[2];
  // MUFU reciprocal: 1 ulp, inside the culling margin.  An axis-parallel direction gets a huge FINITE reciprocal, so that the
  // slab distances can be one FMA each, plane * inv + (-origin * inv), without inf - inf = NaN (with an infinite reciprocal
  // ONE side of a slab became NaN and the other served as both entry and exit — measured: every ray of a collimated source
  // was culled).  The FMA form carries an error of 2^-23 |origin| in position, like the rounding of the origin itself: inside the margin.
  float ix, iy, iz;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"((float)dn[0]));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"((float)dn[1]));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"((float)dn[2]));
  ix = fminf(fmaxf(ix, -1e30f), 1e30f); iy = fminf(fmaxf(iy, -1e30f), 1e30f); iz = fminf(fmaxf(iz, -1e30f), 1e30f);
  const float cx = -sx*ix, cy = -sy*iy, cz = -sz*iz;
  float limf = (float)tmax*1.000002f;                                  // nothing beyond this can still matter
  const bool seq_off = !(FEAT & FEAT_SEQ) || !p.sequential;
  const bool seq_dead = seq_index >= 128;
  const int sw = (seq_index >> 6) & 1, sb = seq_index & 63;
  const int ns = p.scene.n_shells;
  for (int si = 0; si < ns; ++si) {
    const DShell& sh = sshells[si];
    if (si == skip_shell) continue;
    if (!seq_off && (seq_dead || !((sh.seqmask[sw] >> sb) & 1ull))) continue;
    float ta = fmaf(sh.lo[0], ix, cx), tb = fmaf(sh.hi[0], ix, cx);
    float t0 = fminf(ta, tb), t1 = fmaxf(ta, tb);
    ta = fmaf(sh.lo[1], iy, cy); tb = fmaf(sh.hi[1], iy, cy);
    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
    ta = fmaf(sh.lo[2], iz, cz); tb = fmaf(sh.hi[2], iz, cz);
    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
    // miss, entirely behind the start, or beyond what can still matter
    if (t0 > t1 || t1 < 0.0f || t0 > limf) continue;
    const int f1 = sh.face_first + sh.face_count;
    for (int i = sh.face_first; i < f1; ++i) {
      if (sfaces[i].flags & DFACE_PAIR) { test_plane_pair(sfaces[i], sfaces[i+1], i, s, dn, medium, tol, h); ++i; }
      else test_face<false, FEAT>(sfaces[i], i, p, s, dn, medium, seq_index, tmax, h);
    }
    limf = (float)h.lim*1.000002f;
  }
  if (h.fA < 0) return -1;
  if (h.fB >= 0 && h.tB < h.tA + 2*tol) { t_out = h.tB; return h.fB; }   // prefer "not the current medium" (ray.py:445-452)
  t_out = h.tA; return h.fA;
}

// same rule, faces reached through a BVH over face boxes (replaces the shell/face BoundBox culls of ray.py:345-404).
// Conservative fp32 slab tests on boxes widened by the culling margin; near child first, the far child is pushed with
// its entry distance and dropped at pop time if a closer hit has been accepted meanwhile.
template <int FEAT>
__device__ __forceinline__ int find_nearest_bvh(const TraceParams& p, const double* s, const double* dn,
                                                int medium, int seq_index, double max_len, double& t_out) {
  const double tol = p.tol;
  const double tmax = max_len + tol;
  NearestHit h; h.tA = 1e300; h.tB = 1e300; h.lim = tmax; h.fA = -1; h.fB = -1;
  const BvhNode2* __restrict__ nodes = p.scene.bvh;
  const float sx = (float)s[0], sy = (float)s[1], sz = (float)s[2];
  const float ix = __frcp_rn((float)dn[0]), iy = __frcp_rn((float)dn[1]), iz = __frcp_rn((float)dn[2]);
  float limf = (float)tmax*1.000002f;
  int stack_node[ODW_BVH_STACK]; float stack_t[ODW_BVH_STACK]; int sp = 0;
  int node = 0;
  for (;;) {
    const float4* q = reinterpret_cast<const float4*>(nodes + node);
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const int4 d = __ldg(reinterpret_cast<const int4*>(q + 3));
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
    bool hit0 = d.z >= 0 && n0 <= f0 && f0 >= 0.0f && n0 <= limf;
    bool hit1 = d.w >= 0 && n1 <= f1 && f1 >= 0.0f && n1 <= limf;
    if (hit0 && d.z > 0) {                                             // leaf: exact fp64 tests
      for (int k = 0; k < d.z; ++k) {
        const int fi = __ldg(p.scene.bvh_prims + d.x + k);
        test_face<true, FEAT>(p.scene.faces[fi], fi, p, s, dn, medium, seq_index, tmax, h);
      }
      limf = (float)h.lim*1.000002f;
      hit0 = false;
      hit1 = hit1 && n1 <= limf;
    }
    if (hit1 && d.w > 0) {
      for (int k = 0; k < d.w; ++k) {
        const int fi = __ldg(p.scene.bvh_prims + d.y + k);
        test_face<true, FEAT>(p.scene.faces[fi], fi, p, s, dn, medium, seq_index, tmax, h);
      }
      limf = (float)h.lim*1.000002f;
      hit1 = false;
      hit0 = hit0 && n0 <= limf;
    }
    if (hit0 && hit1) {
      const bool first0 = n0 <= n1;
      if (sp < ODW_BVH_STACK) { stack_node[sp] = first0 ? d.y : d.x; stack_t[sp] = first0 ? n1 : n0; ++sp; }
      node = first0 ? d.x : d.y;
      continue;
    }
    if (hit0) { node = d.x; continue; }
    if (hit1) { node = d.y; continue; }
    bool found = false;
    while (sp > 0) {
      --sp;
      if (stack_t[sp] <= limf) { node = stack_node[sp]; found = true; break; }
    }
    if (!found) break;
  }
  if (h.fA < 0) return -1;
  if (h.fB >= 0 && h.tB < h.tA + 2*tol) { t_out = h.tB; return h.fB; }
  t_out = h.tA; return h.fA;
}

// OpticalGroupProxy.onRayHit -> SimulationResults.addRayHit (optical_group.py:206-209, results_store.py:641-648):
// warp-aggregated append (one atomic per converged group of lanes) + optional detector binning
template <int FEAT>
__device__ __forceinline__ void record_hit(const TraceParams& p, unsigned long long ray, int bounce, int group, int face_id,
                                           const double* P, const double* dir, double power, bool entering, int medium,
                                           unsigned int* s_cnt) {
  for (int b = 0; (FEAT & FEAT_BIN) && b < p.n_binnings; ++b) {
    const DBinning& bn = p.binnings[b];
    if (bn.group != group) continue;
    double w[3] = { P[0]-bn.origin[0], P[1]-bn.origin[1], P[2]-bn.origin[2] };
    double x = dot3(w, bn.ua), y = dot3(w, bn.va);
    if (x >= bn.u_lo && x <= bn.u_hi && y >= bn.v_lo && y <= bn.v_hi) {
      int ix = min(bn.nu-1, (int)((x - bn.u_lo)*bn.u_scale));
      int iy = min(bn.nv-1, (int)((y - bn.v_lo)*bn.v_scale));
      atomicAdd(p.bins + bn.offset + (size_t)ix*bn.nv + iy, bn.weighted ? power : 1.0);
    }
  }
  if (!p.store_hits) { atomicAdd(&s_cnt[CNT_HITS], 1u); return; }
  cg::coalesced_group g = cg::coalesced_threads();
  unsigned long long base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&p.counters->hits, (unsigned long long)g.size());
  base = g.shfl(base, 0);
  unsigned long long slot = base + g.thread_rank();
  if (slot >= p.hits.capacity) { atomicAdd(&s_cnt[CNT_DROPPED], 1u); return; }
  double* hp = p.hits.points + 3*slot; hp[0] = P[0]; hp[1] = P[1]; hp[2] = P[2];
  double* hd = p.hits.dirs + 3*slot;   hd[0] = dir[0]; hd[1] = dir[1]; hd[2] = dir[2];
  p.hits.powers[slot] = power;
  p.hits.entering[slot] = entering ? 1 : 0;
  p.hits.ray_index[slot] = ray;
  p.hits.group[slot] = group;
  p.hits.bounce[slot] = bounce;
  p.hits.face_id[slot] = face_id;
  p.hits.medium[slot] = medium;
}

// Philox draw + tabulated inverse CDF + _makeRay: once per ray, kept out of line so the bounce loop stays small
struct RayInit { double o[3], d[3]; };
__device__ __noinline__ RayInit init_ray_mc(const TraceParams& p, unsigned long long ray) {
  double u0, u1, first, phi;
  RayInit r;
  philox_uniform2(p.seed, (uint32_t)p.src.source_id, ray, 0u, u0, u1);
  sample_source(p.src, u0, u1, first, phi);
  make_ray(p.src, first, phi, r.o, r.d);
  return r;
}

// One Monte-Carlo ray of a surface source (SurfaceSourceProxy._generateRays 'true', surface_source.py:522-555): face by
// area weight, area-uniform point redrawn until it lies on the trimmed face, theta from the tabulated density, phi uniform,
// d = cos(theta) n + sin(theta) (cos(phi) (t x n) + sin(phi) t).  Philox purposes: 0 -> (face, theta), 1 -> (phi, -),
// 2+k -> (u, v) of try k, 0x100+k -> acceptance uniform of try k.
#define ODW_SURFACE_MAX_TRIES 64
// first face with a0 < emit_cdf[k]: the guide table brackets it (a tessellated emitter has 1e4..1e5 faces: a plain binary
// search is 17 dependent trips to L2)
__device__ __forceinline__ int pick_emit_face(const DSource& s, double a0) {
  const int cell = min(ODW_EMIT_GUIDE-1, (int)(a0*(double)ODW_EMIT_GUIDE));
  int k = (int)__ldg(s.emit_guide + cell), hi = min(s.n_emit-1, (int)__ldg(s.emit_guide + cell + 1));
  while (k < hi) { const int m = (k + hi) >> 1; if (a0 < __ldg(s.emit_cdf + m)) hi = m; else k = m + 1; }
  return k;
}

// face >= 0: the emitting face of this ray as pick_emit_face found it earlier (sample_kernel regroups rays by face class)
__device__ __noinline__ RayInit init_ray_surface(const DSource& s, unsigned long long seed, unsigned long long ray,
                                                 double* theta_out, double* phi_out, int face = -1) {
  double a0, a1, b0, b1;
  philox_uniform2(seed, (uint32_t)s.source_id, ray, 0u, a0, a1);
  philox_uniform2(seed, (uint32_t)s.source_id, ray, 1u, b0, b1);
  const int k = face >= 0 ? face : pick_emit_face(s, a0);
  const DFace& f = s.emit_faces[k];
  double P[3] = {0, 0, 0}, du[3] = {1, 0, 0}, dv[3] = {0, 1, 0};
  if (f.flags & DFACE_TRI) {
    // a triangle (tessellated emitter): area-uniform point from two uniforms, nothing to reject
    double w0, w1;
    philox_uniform2(seed, (uint32_t)s.source_id, ray, 2u, w0, w1);
    const double sq = sqrt(w0), b1 = sq*(1.0 - w1), b2 = sq*w1, b0 = 1.0 - sq;
    const double u = b0*f.aux[0] + b1*f.aux[2] + b2*f.aux[4], v = b0*f.aux[1] + b1*f.aux[3] + b2*f.aux[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) { P[i] = f.o[i] + u*f.x[i] + v*f.y[i]; du[i] = f.x[i]; dv[i] = f.y[i]; }
  } else
  for (uint32_t tr = 0; tr < ODW_SURFACE_MAX_TRIES; ++tr) {
    double w0, w1, w2, w3, u, v;
    philox_uniform2(seed, (uint32_t)s.source_id, ray, 2u + tr, w0, w1);
    philox_uniform2(seed, (uint32_t)s.source_id, ray, 0x100u + tr, w2, w3);
    const bool ok = surface_draw_uv(f, w0, w1, w2, u, v);
    surface_eval(f, u, v, P, du, dv);
    if (!ok) continue;
    // surface_source.py:399-408; a point drawn inside the parameter window of a face without trim loops is on the face
    if (f.trim != ODW_TRIM_LOOPS || on_trimmed_face(f, s.emit_segs, P, s.dist_tol)) break;
  }
  const double theta = interp_cdf(a1, s.first_cdf, s.first_guide, s.n_first, s.first_lo, s.first_hi, s.n_first_guide);
  const double phi = b0*ODW_TWO_PI;                                              // surface_source.py:544
  double n[3];
  outward_normal(f, P, n);
  const double lu = sqrt(dot3(du, du)), lv = sqrt(dot3(dv, dv));
  const double* t = (lu > 10*s.dist_tol || lu >= lv) ? du : dv;                  // surface_source.py:549
  const double tl = sqrt(dot3(t, t));
  const double th[3] = { t[0]/tl, t[1]/tl, t[2]/tl };
  const double txn[3] = { th[1]*n[2]-th[2]*n[1], th[2]*n[0]-th[0]*n[2], th[0]*n[1]-th[1]*n[0] };
  double st, ct, sp, cp; sincos(theta, &st, &ct); sincos(phi, &sp, &cp);
  RayInit r;
  double d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = ct*n[i] + st*(cp*txn[i] + sp*th[i]);
  const double dl = sqrt(dot3(d, d));
#pragma unroll
  for (int i = 0; i < 3; ++i) { r.o[i] = P[i]; r.d[i] = d[i]/dl; }
  if (theta_out) *theta_out = theta;
  if (phi_out) *phi_out = phi;
  return r;
}


// ------------------------------------------------------------------------------------------
// Ray state of one lane and the steps of Ray.traceRay around findNearestIntersection, shared by the register-resident
// kernel (odw_kernels.cu) and the wavefront kernels (odw_wavefront.cu).
// direction = dn (unit) times dscale: the reference keeps the un-normalised direction (explicit ray lists need not be
// unit, a mirror preserves the length) and reports it with every hit; one scalar instead of a second vector.
// n_isect = segments of the current ray so far.  (References to plain locals of the kernel: as a struct of values the
// state ends up in local memory in kernels that also index a local traversal stack dynamically.)
struct RayState {
  double* point; double* dn; double& dscale; double& power;
  int& medium; int& seq_index; int& n_isect;
  int& skip_shell;     // shell the next segment cannot hit (it starts on that convex shell and points away from it), -1 = none
};

// next ray: Monte-Carlo draw (Philox counter = global ray index) or row i of the explicit list
template <bool MC, int FEAT>
__device__ __forceinline__ void fetch_ray(const TraceParams& p, unsigned long long i, const RayState& r) {
  if (MC) {
    const RayInit q = ((FEAT & FEAT_SURFSRC) && p.src.kind == ODW_SRC_SURFACE) ? init_ray_surface(p.src, p.seed, p.first_ray + i, nullptr, nullptr)
                                                     : init_ray_mc(p, p.first_ray + i);
    r.point[0] = q.o[0]; r.point[1] = q.o[1]; r.point[2] = q.o[2];
    r.dn[0] = q.d[0]; r.dn[1] = q.d[1]; r.dn[2] = q.d[2];
    r.power = 1.0;
  } else {
    const double* o = p.in_origins + 3*i; const double* d = p.in_dirs + 3*i;
    r.point[0] = o[0]; r.point[1] = o[1]; r.point[2] = o[2];
    r.dn[0] = d[0]; r.dn[1] = d[1]; r.dn[2] = d[2];
    r.power = p.in_powers ? p.in_powers[i] : 1.0;
  }
  const double d2 = dot3(r.dn, r.dn), li = fast_rsqrt(d2);
  r.dn[0] *= li; r.dn[1] *= li; r.dn[2] *= li;
  if (!MC) r.dscale = d2*li;            // Monte-Carlo rays start (and stay, to rounding) unit: no length to carry
  r.medium = -1; r.seq_index = 0; r.n_isect = 0; r.skip_shell = -1;
}

// the ray has ended: count its segments, write the per-ray summary of explicit lists.  CNT: shared-memory counters
template <bool MC>
__device__ __forceinline__ void finish_ray(const TraceParams& p, unsigned long long i, const RayState& r, unsigned int* s_cnt) {
  // every find-nearest call yields exactly one segment (a hit or the escape segment), so segments == n_isect
  atomicAdd(&s_cnt[CNT_SEGMENTS], (unsigned int)r.n_isect);
  if (!MC) {
    if (p.out_nseg) p.out_nseg[i] = r.n_isect;
    if (p.out_final_point) { double* q = p.out_final_point + 3*i; q[0] = r.point[0]; q[1] = r.point[1]; q[2] = r.point[2]; }
    if (p.out_final_power) p.out_final_power[i] = r.power;
    if (p.out_final_medium) p.out_final_medium[i] = r.medium;
  }
}

// applyStochasticRayCorrections (optical_group.py:279-323) of a Mirror / Lens hit: `o` holds the ideal outgoing direction
// on entry.  Out of line: surfaces are ideal in every benchmark scene, the bounce loop should not carry this code.
// family: 0 = Mirror hit or entering Lens hit, 1 = leaving Lens hit, -1 = Mirror group (one family); see odw.h odw_scatter.n_tables
__device__ __noinline__ Vec3 apply_scatter(const DScatter* scatters, int main_i, int mod_i, unsigned long long seed, uint32_t source_id,
                                           unsigned long long ray, int bounce, int family, double dx, double dy, double dz,
                                           double nx, double ny, double nz, double ox, double oy, double oz) {
  const double d_in[3] = { dx, dy, dz }, nrm[3] = { nx, ny, nz };
  double o[3] = { ox, oy, oz };
  if (main_i >= 0) {
    double u0, u1, th, ph;
    philox_uniform2(seed, source_id, ray, 0x10000u + 4u*(uint32_t)bounce, u0, u1);
    DScatter t = scatters[main_i];
    if (t.n_tables > 1) {
      // per-hit density: the member of the family nearest to theta_in = angle(direction, normal) (optical_group.py:288)
      const int K = family < 0 ? t.n_tables : t.n_tables/2;
      const double c = fmin(1.0, fmax(-1.0, dot3(d_in, nrm)/sqrt(dot3(d_in, d_in)*dot3(nrm, nrm))));
      const int k = min(K - 1, max(0, (int)(acos(c)/(ODW_TWO_PI/4)*(double)(K - 1) + 0.5)));
      const size_t m = (size_t)((family > 0 ? K : 0) + k);
      t.phi_cdf += m*(size_t)t.n_phi; t.first_cdf += m*(size_t)t.n_rows*(size_t)t.n_first;
      t.phi_guide += m*(size_t)(ODW_GUIDE + 1); t.first_guide += m*(size_t)t.n_rows*(size_t)(ODW_GUIDE + 1);
    }
    sample_source(t, u0, u1, th, ph);
    scatter_direction(nrm, d_in, th, ph, o);
  }
  if (mod_i >= 0) {
    double u0, u1, th, ph;
    const double cur[3] = { o[0], o[1], o[2] };
    philox_uniform2(seed, source_id, ray, 0x10000u + 4u*(uint32_t)bounce + 1u, u0, u1);
    sample_source(scatters[mod_i], u0, u1, th, ph);
    scatter_direction(cur, d_in, th, ph, o);
  }
  Vec3 r; r.x = o[0]; r.y = o[1]; r.z = o[2];
  return r;
}

// Everything Ray.traceRay does after findNearestIntersection returned (ray.py:105-281): escape segment, or move to the
// hit, absorption in the traversed medium, normal, onRayHit, the OpticalType rule.  true = the ray has ended.
template <bool MC, int FEAT>
__device__ __forceinline__ bool interact(const TraceParams& p, const DFace* face_table, const DShell* shells, const DGroup* __restrict__ groups,
                                         int fi, double t, unsigned long long i, const RayState& r, unsigned int* s_cnt) {
  double* point = r.point; double* dn = r.dn;
  if (fi < 0) {                                                                  // ray.py:105-109
    point[0] += dn[0]*p.max_len; point[1] += dn[1]*p.max_len; point[2] += dn[2]*p.max_len;
    atomicAdd(&s_cnt[CNT_ESCAPED], 1u);
    return true;
  }
  const DFace& f = face_table[fi];
  const int fgroup = f.group;
  const DGroup& g = groups[fgroup];
  const int prev_medium = r.medium;
  point[0] += t*dn[0]; point[1] += t*dn[1]; point[2] += t*dn[2];                 // ray.py:117
  if ((FEAT & FEAT_EXT) && r.medium >= 0) {                                      // ray.py:120-125 (multiplicative)
    double L = groups[r.medium].absorption_length;
    if (L == 0) r.power = 0; else if (isfinite(L)) r.power *= exp(-t/L);
  }
  double nrm[3];
  outward_normal(f, point, nrm);
  const bool entering = dot3(dn, nrm) < 0;                                       // ray.py:473-480
  if (entering) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
  if (g.record || p.record_all) {
    const double ds = MC ? 1.0 : r.dscale;
    const double dir[3] = { dn[0]*ds, dn[1]*ds, dn[2]*ds };
    record_hit<FEAT>(p, p.first_ray + i, r.n_isect-1, fgroup, f.face_id, point, dir, r.power, entering, prev_medium, s_cnt);
  }
  double o[3] = { dn[0], dn[1], dn[2] };                                         // outgoing direction / its length
  double oscale = MC ? 1.0 : r.dscale;
  switch (g.type) {
    case ODW_OPT_MIRROR: {                                                       // ray.py:146-161
      mirror_dir(dn, nrm, o);                                                    // d - 2(d.n)n is linear in d: the length carries over
      if ((FEAT & FEAT_EXT) && p.scene.scatters && (g.scat_main >= 0 || g.scat_modify >= 0)) {   // ray.py:151-155 (uniform test first: no scene table, no loads)
        const Vec3 q = apply_scatter(p.scene.scatters, g.scat_main, g.scat_modify, p.seed, (uint32_t)p.src.source_id, p.first_ray + i,
                                     r.n_isect-1, -1, dn[0], dn[1], dn[2], nrm[0], nrm[1], nrm[2], o[0], o[1], o[2]);
        o[0] = q.x; o[1] = q.y; o[2] = q.z; oscale = 1.0;
      }
      r.power *= g.reflectivity; ++r.seq_index;
      break;
    }
    case ODW_OPT_LENS: {                                                         // ray.py:165-211
      double n1 = r.medium >= 0 ? groups[r.medium].n : 1.0, n2 = 1.0;
      if (entering) n2 = g.n;
      if ((FEAT & FEAT_EXT) && g.fresnel) {                                      // opt-in extension (odw.h odw_group.fresnel): not in the reference
        double u0, u1;
        philox_uniform2(p.seed, (uint32_t)p.src.source_id, p.first_ray + i, 0x20000u + (uint32_t)(r.n_isect-1), u0, u1);
        if (u0 < fresnel_reflectance(dn, nrm, n1, n2)) { mirror_dir(dn, nrm, o); oscale = 1.0; break; }   // reflected: medium and sequence index stay
      }
      if (entering) r.medium = fgroup;
      bool tir = snell(dn, n1, n2, nrm, o);
      oscale = 1.0;                                                              // snellsLaw works on the unit direction
      if ((FEAT & FEAT_EXT) && p.scene.scatters && (g.scat_main >= 0 || g.scat_modify >= 0)) {   // ray.py:197-201
        const Vec3 q = apply_scatter(p.scene.scatters, g.scat_main, g.scat_modify, p.seed, (uint32_t)p.src.source_id, p.first_ray + i,
                                     r.n_isect-1, entering ? 0 : 1, dn[0], dn[1], dn[2], nrm[0], nrm[1], nrm[2], o[0], o[1], o[2]);
        o[0] = q.x; o[1] = q.y; o[2] = q.z;
      }
      if (!entering && !tir && r.medium == fgroup) { r.medium = -1; ++r.seq_index; }
      break;
    }
    case ODW_OPT_GRATING: {                                                      // ray.py:216-268
      if (!(FEAT & FEAT_EXT)) break;                                             // never reached: the host picks an instance with FEAT_EXT for scenes with gratings
      if (g.gtype == ODW_GRATING_REFLECTION) {
        if (entering) {
          double n = r.medium >= 0 ? groups[r.medium].n : 1.0;
          const Vec3 q = line_grating(dn[0], dn[1], dn[2], n, n, nrm[0], nrm[1], nrm[2], &g, p.wavelength, false);
          o[0] = q.x; o[1] = q.y; o[2] = q.z; oscale = 1.0; ++r.seq_index;
        }
      } else if (entering) {
        if (r.medium >= 0) { r.power = 0; break; }                               // the reference raises ValueError here
        r.medium = fgroup;
        const Vec3 q = line_grating(dn[0], dn[1], dn[2], 1.0, g.n, nrm[0], nrm[1], nrm[2], &g, p.wavelength, true);
        o[0] = q.x; o[1] = q.y; o[2] = q.z; oscale = 1.0;
      } else {
        double n1 = r.medium >= 0 ? groups[r.medium].n : 1.0;
        bool tir = snell(dn, n1, 1.0, nrm, o);
        oscale = 1.0;
        if (!tir) { r.medium = -1; ++r.seq_index; }
      }
      break;
    }
    case ODW_OPT_ABSORBER: r.power = 0; ++r.seq_index; break;                    // ray.py:271-273
    default: ++r.seq_index; break;                                               // Vacuum, ray.py:276-277
  }
  if (r.power < p.power_tol) return true;                                        // ray.py:280
  // A ray that leaves the surface of a convex solid outwards cannot meet that solid again: its shell is skipped on the
  // next segment (the reference tests it and finds nothing beyond distTol).  n_out = entering ? -nrm : nrm.
  r.skip_shell = -1;
  if (shells && shells[f.shell].convex) {
    const double out = dot3(o, nrm);
    if (entering ? out < 0 : out > 0) r.skip_shell = f.shell;
  }
  // next segment: unit direction and the length the reference would carry along
  const double o2 = dot3(o, o), li = fast_rsqrt(o2);
  dn[0] = o[0]*li; dn[1] = o[1]*li; dn[2] = o[2]*li;
  if (!MC) r.dscale = oscale*(o2*li);
  return false;
}

}  // namespace
