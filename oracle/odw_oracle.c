/*
 * odw_oracle.c — CPU restatement of the reference's trace loop.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (libodw_b200.so) never links or calls it.
 *
 * PARITY STATUS
 *   pinned   — the whole per-ray path, against the reference's OWN Python executed here with NO method overridden
 *              (tests/golden/make_traceray_golden.py -> tests/golden/traceray_golden.npz, tests/test_traceray_golden.py):
 *              Ray.traceRay (state machine, power / medium / sequence-index bookkeeping, powerTol and maxIntersections
 *              exits), Ray.findNearestIntersection (shell and face candidates by enlarged-box distance, the LINE test of
 *              the boxes, the three acceptance rules, the maxRayLength shrink, the minDist + 2 tol filter, the "not the
 *              current medium" preference), Ray.getNormal (flip / isEntering), mirror, snellsLaw incl. total reflection,
 *              lineGrating, find.relevantOpticalObjects + getTracingSequence (ignore list, sequential filter),
 *              raytracing_cache, OpticalGroupProxy.onRayHit, the rotation formula of applyStochasticRayCorrections and
 *              PointSourceProxy._makeRay.  FreeCAD's Base types and the OpenCASCADE PRIMITIVES the loop calls (line x
 *              untrimmed surface, point-to-edge and point-to-trimmed-face distance, bounding boxes, Surface.parameter,
 *              normalAt) are numpy stand-ins (tests/freecad_stub.py, tests/occ_stub.py) that do not use this file;
 *            — the sampler (numeric mode), the fan grid and the fan ray list: against the importable reference
 *              `distributions` / `point_source` modules (tests/golden/make_sampler_golden.py, make_fan_golden.py).
 *   "parity unpinned" — OpenCASCADE's primitive answers on real BRep shapes only: the reference delegates them to
 *              FreeCAD/OCC (unpinned "system FreeCAD / latest AppImage", benchmark files written by FreeCAD 1.1R20260725),
 *              absent here AND on the GPU box (profiles/r02_freecad_probe.txt), and its tests hold no golden vector for
 *              a single ray.  The stand-ins restate their documented behaviour with a different algorithm (polynomial
 *              root finding); hand-derived known answers anchor them (tests/test_occ_stub.py,
 *              tests/test_oracle_known_answers.py, tests/test_conicoid.py).
 *
 * Each function cites the reference code it follows (paths relative to
 * /root/reference/freecad/optics_design_workbench/).  Where the reference calls OCC
 * (Curve.intersect(Surface), distToShape, normalAt) the closed-form equivalent for
 * plane / cylinder / cone / sphere / torus / conic of revolution (+ even-asphere terms) is written out.
 *
 * Plain scalar C, one ray at a time, same loop structure as the reference:
 * groups -> shells (bbox cull, sorted) -> faces (bbox cull, sorted) -> all line/surface points ->
 * the three acceptance rules -> maxRayLength shrink -> final choice.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/odw.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define TWO_PI 6.283185307179586476925286766559

/* ------------------------------------------------------------------------------------------ */
/* small vector helpers                                                                       */

static inline double dot3(const double* a, const double* b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
static inline void cross3(const double* a, const double* b, double* o) {
  o[0]=a[1]*b[2]-a[2]*b[1]; o[1]=a[2]*b[0]-a[0]*b[2]; o[2]=a[0]*b[1]-a[1]*b[0];
}
static inline double len3(const double* a) { return sqrt(dot3(a,a)); }

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), counter = (ray_lo, ray_hi, source_id, purpose), key = seed */

static inline void philox_round(uint32_t* c, const uint32_t* k) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  uint32_t n0 = hi1 ^ c[1] ^ k[0];
  uint32_t n1 = lo1;
  uint32_t n2 = hi0 ^ c[3] ^ k[1];
  uint32_t n3 = lo0;
  c[0]=n0; c[1]=n1; c[2]=n2; c[3]=n3;
}

void oracle_philox(uint64_t seed, uint32_t source_id, uint64_t ray, uint32_t purpose, double* u2) {
  uint32_t c[4] = { (uint32_t)ray, (uint32_t)(ray >> 32), source_id, purpose };
  uint32_t k[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
  uint64_t a = ((uint64_t)c[0] << 32) | c[1];
  uint64_t b = ((uint64_t)c[2] << 32) | c[3];
  u2[0] = (double)(a >> 11) * (1.0/9007199254740992.0);   /* [0,1) with 53 bits, like numpy random_sample */
  u2[1] = (double)(b >> 11) * (1.0/9007199254740992.0);
}

/* ------------------------------------------------------------------------------------------ */
/* sampler: distributions/random_number_generator.py:413-456 (interpolateResult) + :492-500 (draw) */

/* numpy.linspace(lo, hi, n)[i]  (start + i*step, last element forced to hi) */
static inline double linspace_at(double lo, double hi, int n, int i) {
  if (i >= n-1) return hi;
  return lo + (double)i * ((hi-lo)/(double)(n-1));
}

/* numpy.interp(x, xp, fp) with xp increasing (repeats allowed), fp = linspace(lo,hi,n) */
static double interp_cdf(double x, const double* xp, int n, double lo, double hi) {
  if (x > xp[n-1]) return hi;
  if (x < xp[0]) return lo;
  /* j = last index with xp[j] <= x  (binary_search_with_guess semantics) */
  int a = 0, b = n;           /* invariant: xp[a] <= x, (b==n or xp[b] > x) */
  while (b - a > 1) {
    int m = (a + b) >> 1;
    if (xp[m] <= x) a = m; else b = m;
  }
  int j = a;
  if (j == n-1) return hi;
  double fj = linspace_at(lo, hi, n, j);
  if (xp[j] == x) return fj;
  double fj1 = linspace_at(lo, hi, n, j+1);
  double slope = (fj1 - fj) / (xp[j+1] - xp[j]);
  return slope * (x - xp[j]) + fj;
}

/* argmin |C - phi| over mid-points C[i] = (E[i+1]+E[i])/2, first minimum wins (numpy.argmin) */
static int nearest_row(double phi, double lo, double hi, int n_edges) {
  int nrows = n_edges - 1;
  double step = (hi-lo)/(double)(n_edges-1);
  int k = (int)floor((phi - lo)/step);
  if (k < 0) k = 0;
  if (k > nrows-1) k = nrows-1;
  int best = -1; double bestd = 0;
  for (int i = k-2; i <= k+2; ++i) {
    if (i < 0 || i >= nrows) continue;
    double c = (linspace_at(lo,hi,n_edges,i+1) + linspace_at(lo,hi,n_edges,i))/2;
    double d = fabs(c - phi);
    if (best < 0 || d < bestd) { best = i; bestd = d; }
  }
  return best;
}

/* stochastic surface model: (theta, phi) from a tabulated density, same draw as the point source sampler */
static void scatter_sample(const odw_scatter* s, double u_phi, double u_first, double* theta, double* phi) {
  double ph = interp_cdf(u_phi, s->phi_cdf, s->n_phi, s->phi_lo, s->phi_hi);
  int row = 0;
  if (s->n_rows > 1) row = nearest_row(ph, s->phi_lo, s->phi_hi, s->n_phi);
  *theta = interp_cdf(u_first, s->first_cdf + (size_t)row*(size_t)s->n_first, s->n_first, s->first_lo, s->first_hi);
  *phi = ph;
}

/* Rotation(n, phi) * Rotation(n x d_in, theta) * n  (optical_group.py:309-311,318-320):
 * out = cos(theta) n^ + sin(theta) (cos(phi) (a x n^) + sin(phi) a),  a = unit(n x d_in); a zero axis rotates nothing */
static void scatter_direction(const double* n, const double* d_in, double theta, double phi, double* out) {
  double nl = len3(n), nh[3] = { n[0]/nl, n[1]/nl, n[2]/nl }, a[3];
  cross3(nh, d_in, a);
  double al = len3(a);
  if (!(al > 1e-300)) { for (int i = 0; i < 3; ++i) out[i] = nh[i]*nl; return; }
  for (int i = 0; i < 3; ++i) a[i] /= al;
  double axn[3]; cross3(a, nh, axn);
  double st = sin(theta), ct = cos(theta), sp = sin(phi), cp = cos(phi);
  for (int i = 0; i < 3; ++i) out[i] = nl*(ct*nh[i] + st*(cp*axn[i] + sp*a[i]));
}

/* applyStochasticRayCorrections (optical_group.py:279-323) for a Mirror / Lens hit; dir_in unit, out = ideal direction on entry */
/* member of a per-hit family (odw.h odw_scatter.n_tables > 1) for a hit: the table nearest to theta_in = angle(direction,
 * normal), optical_group.py:288; Lens groups carry two families (entering, leaving) */
static int scatter_member(const odw_scatter* t, const double* dir_in, const double* nrm, int lens, int entering) {
  if (t->n_tables <= 1) return 0;
  int K = lens ? t->n_tables/2 : t->n_tables;
  double c = dot3(dir_in, nrm)/(len3(dir_in)*len3(nrm));
  if (c > 1) c = 1; if (c < -1) c = -1;
  double th = acos(c);
  int k = (int)(th/(TWO_PI/4)*(double)(K - 1) + 0.5);
  if (k < 0) k = 0; if (k > K - 1) k = K - 1;
  return (lens && !entering ? K : 0) + k;
}

static void apply_scatter(const odw_scene_desc* sc, int group, uint64_t seed, uint32_t source_id, uint64_t ray, int bounce,
                          const double* dir_in, const double* nrm, int entering, double* out) {
  if (!sc->n_scatters || !sc->group_scatter) return;
  int main_i = sc->group_scatter[2*group], mod_i = sc->group_scatter[2*group+1];
  if (main_i >= 0) {
    double u[2], th, ph;
    oracle_philox(seed, source_id, ray, 0x10000u + 4u*(uint32_t)bounce, u);
    odw_scatter t = sc->scatters[main_i];
    int m = scatter_member(&t, dir_in, nrm, sc->groups[group].optical_type == ODW_OPT_LENS, entering);
    t.phi_cdf += (size_t)m*(size_t)t.n_phi; t.first_cdf += (size_t)m*(size_t)t.n_rows*(size_t)t.n_first;
    scatter_sample(&t, u[0], u[1], &th, &ph);
    scatter_direction(nrm, dir_in, th, ph, out);
  }
  if (mod_i >= 0) {
    double u[2], th, ph, cur[3] = { out[0], out[1], out[2] };
    oracle_philox(seed, source_id, ray, 0x10000u + 4u*(uint32_t)bounce + 1u, u);
    scatter_sample(&sc->scatters[mod_i], u[0], u[1], &th, &ph);
    scatter_direction(cur, dir_in, th, ph, out);
  }
}

static void sample_from_uniforms(const odw_source_desc* s, double u_phi, double u_first,
                                 double* first, double* phi) {
  /* phi (last variable) is drawn first from its marginal; then theta|r conditional on the nearest phi row */
  double ph = interp_cdf(u_phi, s->phi_cdf, s->n_phi, s->phi_lo, s->phi_hi);
  int row = 0;
  if (s->n_rows > 1) row = nearest_row(ph, s->phi_lo, s->phi_hi, s->n_phi);
  const double* cdf = s->first_cdf + (size_t)row * (size_t)s->n_first;
  *first = interp_cdf(u_first, cdf, s->n_first, s->first_lo, s->first_hi);
  *phi = ph;
}

/* freecad_elements/point_source.py:411-460 (_makeRay) */
static void make_ray(const odw_source_desc* s, double first, double phi, double* origin, double* dir) {
  double lo[3], ld[3];
  if (s->kind == ODW_SRC_POINT_SPHERICAL) {
    double st = sin(first), ct = cos(first), sp = sin(phi), cp = cos(phi);
    ld[0] = st*sp; ld[1] = -st*cp; ld[2] = ct;          /* Rz(phi) * Rx(theta) * (0,0,1) */
    double f = s->focal_length;
    lo[0] = (0.0 - ld[0])*f; lo[1] = (0.0 - ld[1])*f; lo[2] = (1.0 - ld[2])*f;
  } else {
    ld[0] = 0; ld[1] = 0; ld[2] = 1;
    lo[0] = first*cos(phi); lo[1] = -first*sin(phi); lo[2] = 0;   /* r*x^*cos + r*(x^ x z^)*sin */
  }
  double l = len3(ld);
  double p2l[3] = { lo[0]+ld[0]/l, lo[1]+ld[1]/l, lo[2]+ld[2]/l };
  const double* M = s->gpM;
  double p1[3], p2[3];
  for (int i = 0; i < 3; ++i) {
    p1[i] = M[4*i+0]*lo[0]  + M[4*i+1]*lo[1]  + M[4*i+2]*lo[2]  + M[4*i+3];
    p2[i] = M[4*i+0]*p2l[0] + M[4*i+1]*p2l[1] + M[4*i+2]*p2l[2] + M[4*i+3];
  }
  double d[3] = { p2[0]-p1[0], p2[1]-p1[1], p2[2]-p1[2] };
  double dl = len3(d);
  for (int i = 0; i < 3; ++i) { origin[i] = p1[i]; dir[i] = d[i]/dl; }
}

int oracle_sample_uniforms(const odw_source_desc* s, const double* u_phi, const double* u_first,
                           uint64_t n, double* first, double* phi) {
  for (uint64_t i = 0; i < n; ++i) sample_from_uniforms(s, u_phi[i], u_first[i], &first[i], &phi[i]);
  return 0;
}

static void source_make_ray(const odw_source_desc* s, uint64_t seed, uint64_t ray, double* first, double* phi,
                            double* origin, double* dir);

int oracle_sample_mc(const odw_source_desc* s, uint64_t seed, uint64_t first_ray, uint64_t n,
                     double* first, double* phi, double* origins, double* dirs) {
  for (uint64_t i = 0; i < n; ++i) {
    double f, p, o[3], d[3];
    source_make_ray(s, seed, first_ray + i, &f, &p, o, d);
    if (first) first[i] = f;
    if (phi) phi[i] = p;
    if (origins) memcpy(origins + 3*i, o, sizeof o);
    if (dirs) memcpy(dirs + 3*i, d, sizeof d);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* line / untrimmed-surface intersection  (ray.py:411  line.Curve.intersect(Surface), infinite line) */

static int solve_quadratic(double a, double b, double c, double* t) {
  /* a t^2 + b t + c = 0, real roots, numerically stable */
  if (a == 0) {
    if (b == 0) return 0;
    t[0] = -c/b; return 1;
  }
  double disc = b*b - 4*a*c;
  if (disc < 0) return 0;
  double sq = sqrt(disc);
  double q = -0.5*(b + (b >= 0 ? sq : -sq));
  int n = 0;
  t[n++] = q/a;
  if (q != 0) t[n++] = c/q; else t[n++] = 0.0;
  return n;
}

static int solve_cubic_depressed(double p, double q, double* y) {
  /* y^3 + p y + q = 0 */
  int n = 0;
  double disc = q*q/4 + p*p*p/27;
  if (disc > 0) {
    double sq = sqrt(disc);
    double A = cbrt(-q/2 + sq), B = cbrt(-q/2 - sq);
    y[n++] = A + B;
  } else if (p == 0) {
    y[n++] = cbrt(-q);
  } else {
    double m = 2*sqrt(-p/3);
    double arg = 3*q/(p*m);
    if (arg > 1) arg = 1; if (arg < -1) arg = -1;
    double th = acos(arg)/3;
    for (int k = 0; k < 3; ++k) y[n++] = m*cos(th - TWO_PI*k/3);
  }
  for (int i = 0; i < n; ++i)            /* Newton polish */
    for (int it = 0; it < 3; ++it) {
      double f = (y[i]*y[i] + p)*y[i] + q, df = 3*y[i]*y[i] + p;
      if (df != 0) y[i] -= f/df;
    }
  return n;
}

/* real roots of the depressed quartic t^4 + B t^2 + C t + D in [lo, hi]: bracket between the critical
 * points (roots of 4t^3 + 2Bt + C), then safeguarded Newton in every monotone piece with a sign change */
static int solve_quartic_depressed(double B, double C, double D, double lo, double hi, double* roots) {
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

/* ---- even asphere on top of a conic of revolution (ODW_SURF_CONICOID with an ODW_SEG_ASPHERE record, odw.h) ----
 * sag as a function of u = rho^2:  S(u) = c u / (1 + sqrt(1 - (1+k) c^2 u)) + u^2 (a0 + a1 u + a2 u^2 + a3 u^3 + a4 u^4) */
static const double* asphere_coeffs(const odw_face* f, const odw_trimseg* segs) {
  if (f->kind != ODW_SURF_CONICOID || f->seg_count < 1 || !segs) return NULL;
  return segs[f->seg_first].kind == ODW_SEG_ASPHERE ? segs[f->seg_first].a : NULL;
}

static int asphere_sag(double c, double k, const double* a, double u, double* S, double* dSdu) {
  double q2 = 1.0 - (1.0 + k)*c*c*u;
  if (!(q2 > 1e-14)) return 0;                              /* beyond (or on) the equator of the base conic */
  double q = sqrt(q2);
  *S = c*u/(1.0 + q) + u*u*(a[0] + u*(a[1] + u*(a[2] + u*(a[3] + u*a[4]))));
  *dSdu = c/(2.0*q) + u*(2*a[0] + u*(3*a[1] + u*(4*a[2] + u*(5*a[3] + u*6*a[4]))));
  return 1;
}

/* Newton on g(t) = z(t) - S(u(t)) along w + t d (d unit), started at t0; 1 = converged onto the surface */
static int asphere_newton(double c, double k, const double* a, const double* w, const double* d, double wz, double dz,
                          double t0, double* t_out) {
  double wd = dot3(w, d), ww = dot3(w, w), t = t0;
  for (int it = 0; it < 40; ++it) {
    double z = wz + t*dz, u = ww + t*(2*wd + t) - z*z, S, dS;
    if (u < 0) u = 0;
    if (!asphere_sag(c, k, a, u, &S, &dS)) return 0;
    double g = z - S, gp = dz - dS*2*((wd + t) - z*dz);
    if (gp == 0 || !isfinite(gp)) return 0;
    double dt = g/gp;
    t -= dt;
    if (fabs(dt) <= 1e-15*fmax(1.0, fabs(t))) {
      z = wz + t*dz; u = ww + t*(2*wd + t) - z*z; if (u < 0) u = 0;
      if (!asphere_sag(c, k, a, u, &S, &dS)) return 0;
      if (fabs(z - S) > 1e-10*(1.0 + fabs(z))) return 0;
      *t_out = t; return 1;
    }
  }
  return 0;
}

/* all parameters t with start + t*d on the untrimmed surface; d must be unit.  returns count (<=4) */
static int line_surface(const odw_face* f, const odw_trimseg* segs, const double* s, const double* d, double* t) {
  double w[3] = { s[0]-f->origin[0], s[1]-f->origin[1], s[2]-f->origin[2] };
  switch (f->kind) {
    case ODW_SURF_PLANE: {
      double den = dot3(d, f->zdir);
      if (den == 0) return 0;
      t[0] = -dot3(w, f->zdir)/den;
      return 1;
    }
    case ODW_SURF_SPHERE: {
      double b = 2*dot3(w, d), c = dot3(w, w) - f->p0*f->p0;
      return solve_quadratic(1.0, b, c, t);
    }
    case ODW_SURF_CYLINDER: {
      double wz = dot3(w, f->zdir), dz = dot3(d, f->zdir);
      double a = 1.0 - dz*dz, b = 2*(dot3(w, d) - wz*dz), c = dot3(w, w) - wz*wz - f->p0*f->p0;
      if (fabs(a) < 1e-300) return 0;
      return solve_quadratic(a, b, c, t);
    }
    case ODW_SURF_CONE: {
      /* |w_perp|^2 = (r + h tan(a))^2, h = axial coordinate */
      double ta = tan(f->p1);
      double wz = dot3(w, f->zdir), dz = dot3(d, f->zdir);
      double r0 = f->p0 + wz*ta, r1 = dz*ta;
      double a = 1.0 - dz*dz - r1*r1;
      double b = 2*(dot3(w, d) - wz*dz - r0*r1);
      double c = dot3(w, w) - wz*wz - r0*r0;
      return solve_quadratic(a, b, c, t);
    }
    case ODW_SURF_TORUS: {
      double R = f->p0, r = f->p1;
      double t0 = -dot3(w, d);                       /* shift to the point of closest approach to the centre */
      double o[3] = { w[0]+t0*d[0], w[1]+t0*d[1], w[2]+t0*d[2] };
      double m = dot3(o, o);
      double rr = (R + r)*(R + r);
      if (m > rr) return 0;
      double half = sqrt(rr - m) + 1e-9;
      double oz = dot3(o, f->zdir), dz = dot3(d, f->zdir);
      double oxy2 = m - oz*oz, dxy2 = 1.0 - dz*dz, g = -oz*dz;   /* o.d = 0  =>  o_xy.d_xy = -oz*dz */
      double K = m + R*R - r*r;
      double B = 2*K - 4*R*R*dxy2, C = -8*R*R*g, D = K*K - 4*R*R*oxy2;
      double roots[4];
      int n = solve_quartic_depressed(B, C, D, -half, half, roots);
      for (int i = 0; i < n; ++i) t[i] = t0 + roots[i];
      return n;
    }
    case ODW_SURF_CONICOID: {
      /* c (rho^2 + (1+k) z^2) - 2 z = 0 along w + t d:  A t^2 + B t + C = 0; a paraboloid met along its axis has A = 0 */
      double c = f->p0, k = f->p1;
      double wz = dot3(w, f->zdir), dz = dot3(d, f->zdir);
      double A = c*(1.0 + k*dz*dz), B = 2*(c*(dot3(w, d) + k*wz*dz) - dz), C = c*(dot3(w, w) + k*wz*wz) - 2*wz;
      double r[2];
      int n = solve_quadratic(A, B, C, r), m = 0;
      /* the sheet of the quadric that the sag formula describes: q = 1 - (1+k) c z >= 0 */
      for (int i = 0; i < n; ++i) if (1.0 - (1.0 + k)*c*(wz + r[i]*dz) >= 0) t[m++] = r[i];
      const double* a = asphere_coeffs(f, segs);
      if (a) {
        /* polynomial terms: Newton from every crossing of the base conic (from the vertex plane when it has none) */
        double start[2]; int ns = m;
        for (int i = 0; i < m; ++i) start[i] = t[i];
        if (ns == 0 && dz != 0) { start[0] = -wz/dz; ns = 1; }
        m = 0;
        for (int i = 0; i < ns; ++i) {
          double tt;
          if (!asphere_newton(c, k, a, w, d, wz, dz, start[i], &tt)) continue;
          if (m == 1 && fabs(tt - t[0]) <= 1e-9*fmax(1.0, fabs(tt))) continue;    /* both starts found the same crossing */
          t[m++] = tt;
        }
      }
      return m;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* (u, v) of a point on the surface and the outward normal  (ray.py:463-465 Surface.parameter + normalAt) */

static void surface_uv_normal(const odw_face* f, const odw_trimseg* segs, const double* P, double* uv, double* n_out) {
  double w[3] = { P[0]-f->origin[0], P[1]-f->origin[1], P[2]-f->origin[2] };
  double x = dot3(w, f->xdir), y = dot3(w, f->ydir), z = dot3(w, f->zdir);
  double ng[3];
  switch (f->kind) {
    case ODW_SURF_PLANE:
      uv[0] = x; uv[1] = y;
      ng[0] = f->zdir[0]; ng[1] = f->zdir[1]; ng[2] = f->zdir[2];
      break;
    case ODW_SURF_CYLINDER: {
      uv[0] = atan2(y, x); uv[1] = z;
      double rho = sqrt(x*x + y*y);
      for (int i = 0; i < 3; ++i) ng[i] = (x*f->xdir[i] + y*f->ydir[i])/rho;
      break;
    }
    case ODW_SURF_CONE: {
      double ca = cos(f->p1), sa = sin(f->p1);
      double v = z/ca;
      double rho_signed = f->p0 + v*sa;
      double rho = sqrt(x*x + y*y);
      /* on the second nappe the radial direction of the parametrisation points the other way */
      double sgn = rho_signed >= 0 ? 1.0 : -1.0;
      uv[0] = atan2(sgn*y, sgn*x); uv[1] = v;
      for (int i = 0; i < 3; ++i)
        ng[i] = sgn*(ca*(x*f->xdir[i] + y*f->ydir[i])/rho*sgn - sa*f->zdir[i]);
      break;
    }
    case ODW_SURF_SPHERE: {
      double R = f->p0;
      double sv = z/R; if (sv > 1) sv = 1; if (sv < -1) sv = -1;
      uv[0] = atan2(y, x); uv[1] = asin(sv);
      for (int i = 0; i < 3; ++i) ng[i] = w[i]/R;
      break;
    }
    case ODW_SURF_TORUS: {
      double R = f->p0, r = f->p1;
      double rho = sqrt(x*x + y*y);
      uv[0] = atan2(y, x); uv[1] = atan2(z, rho - R);
      for (int i = 0; i < 3; ++i) {
        double ring = R*(x*f->xdir[i] + y*f->ydir[i])/rho;   /* centre of the tube circle */
        ng[i] = (w[i] - ring)/r;
      }
      break;
    }
    case ODW_SURF_CONICOID: {
      /* v = rho; du x dv is along c rho_vec - q Z, q = 1 - (1+k) c z (= the square root of the sag formula) */
      double q = 1.0 - (1.0 + f->p1)*f->p0*z;
      uv[0] = atan2(y, x); uv[1] = sqrt(x*x + y*y);
      const double* a = asphere_coeffs(f, segs);
      if (a) {       /* S_rho e_r - Z with S_rho = 2 rho dS/du: the same direction as c rho_vec - q Z when the polynomial vanishes */
        double S, dS;
        if (!asphere_sag(f->p0, f->p1, a, x*x + y*y, &S, &dS)) { S = 0; dS = 0; }
        for (int i = 0; i < 3; ++i) ng[i] = 2*dS*(x*f->xdir[i] + y*f->ydir[i]) - f->zdir[i];
        break;
      }
      for (int i = 0; i < 3; ++i) ng[i] = f->p0*(x*f->xdir[i] + y*f->ydir[i]) - q*f->zdir[i];
      break;
    }
    default:
      uv[0] = uv[1] = 0; ng[0] = ng[1] = 0; ng[2] = 1;
  }
  double l = len3(ng);
  for (int i = 0; i < 3; ++i) n_out[i] = (double)f->nsign * ng[i]/l;
}

/* ------------------------------------------------------------------------------------------ */
/* "vert.distToShape(face) < distTol" (ray.py:426) in (u,v) space: inside the trim region dilated by the tolerance */

static double wrap_into(double x, double lo) {      /* smallest x + 2*pi*k >= lo */
  double y = x - TWO_PI*floor((x - lo)/TWO_PI);
  return y;
}

static double seg_dist_line(double u, double v, const double* a, double su, double sv) {
  double ax = a[0]*su, ay = a[1]*sv, bx = a[2]*su, by = a[3]*sv, px = u*su, py = v*sv;
  double dx = bx-ax, dy = by-ay;
  double l2 = dx*dx + dy*dy;
  double t = l2 > 0 ? ((px-ax)*dx + (py-ay)*dy)/l2 : 0;
  if (t < 0) t = 0; if (t > 1) t = 1;
  double qx = ax + t*dx - px, qy = ay + t*dy - py;
  return sqrt(qx*qx + qy*qy);
}

static int trim_contains(const odw_face* f, const odw_trimseg* segs, double* uv, double tol) {
  if (f->trim_kind == ODW_TRIM_NONE) return 1;
  double u = uv[0], v = uv[1];
  /* metric of the parametrisation at this point: ds = su*du, sv*dv */
  double su = 1, sv = 1;
  int uper = 0, vper = 0;
  switch (f->kind) {
    case ODW_SURF_CYLINDER: su = f->p0; uper = 1; break;
    case ODW_SURF_CONE:     su = fabs(f->p0 + v*sin(f->p1)); uper = 1; break;
    case ODW_SURF_SPHERE:   su = f->p0*cos(v); sv = f->p0; uper = 1; break;
    case ODW_SURF_TORUS:    su = f->p0 + f->p1*cos(v); sv = f->p1; uper = 1; vper = 1; break;
    case ODW_SURF_CONICOID: {   /* meridian arc length per unit rho: sqrt(1 + z'^2), z' = c rho / q */
      double q2 = 1.0 - (1.0 + f->p1)*f->p0*f->p0*v*v;
      if (q2 < 1e-12) q2 = 1e-12;
      double slope = f->p0*v/sqrt(q2);
      const double* a = asphere_coeffs(f, segs);
      if (a) { double u2 = v*v; slope += 2*v*u2*(2*a[0] + u2*(3*a[1] + u2*(4*a[2] + u2*(5*a[3] + u2*6*a[4])))); }
      su = v; sv = sqrt(1.0 + slope*slope); uper = 1; break;
    }
  }
  if (su < 1e-12) su = 1e-12;
  double tu = tol/su, tv = tol/sv;
  if (uper) u = wrap_into(u, f->uv_min[0] - tu);
  if (vper) v = wrap_into(v, f->uv_min[1] - tv);
  uv[0] = u; uv[1] = v;
  if (u < f->uv_min[0] - tu || u > f->uv_max[0] + tu || v < f->uv_min[1] - tv || v > f->uv_max[1] + tv) return 0;
  if (f->trim_kind == ODW_TRIM_UVBOX) return 1;
  /* even-odd crossing count of the half line u' > u, plus closeness to the boundary */
  int crossings = 0;
  for (int i = 0; i < f->seg_count; ++i) {
    const odw_trimseg* s = &segs[f->seg_first + i];
    const double* a = s->a;
    if (s->kind == ODW_SEG_ASPHERE) continue;          /* auxiliary record, not a boundary piece */
    if (s->kind == ODW_SEG_LINE) {
      if ((a[1] > v) != (a[3] > v)) {
        double ux = a[0] + (v - a[1])*(a[2] - a[0])/(a[3] - a[1]);
        if (ux > u) ++crossings;
      }
      if (seg_dist_line(u, v, a, su, sv) < tol) return 1;
    } else {
      double cu = a[0], cv = a[1], r = a[2], a0 = a[3], span = a[4];
      double dv = v - cv;
      if (fabs(dv) < r) {
        double h = sqrt(r*r - dv*dv);
        for (int k = 0; k < 2; ++k) {
          double ux = k ? cu + h : cu - h;
          if (ux > u) {
            double ang = atan2(dv, ux - cu);
            double rel = ang - a0; rel -= TWO_PI*floor(rel/TWO_PI);
            if (rel <= span) ++crossings;
          }
        }
      }
      /* distance to the arc (isotropic metric assumed: arcs only occur on planes / as polylines elsewhere) */
      double du = u - cu, rad = sqrt(du*du + dv*dv);
      double ang = atan2(dv, du), rel = ang - a0; rel -= TWO_PI*floor(rel/TWO_PI);
      if (rel <= span && fabs(rad - r)*fmin(su, sv) < tol) return 1;
    }
  }
  return crossings & 1;
}

/* ------------------------------------------------------------------------------------------ */
/* surface source: freecad_elements/surface_source.py:418-555 (_generateRays 'true'), :390-410
 * (_drawRandomPositionOnFace), :85-111 (_makeRay).  The reference draws (u,v) from a tabulated area element
 * (:269-387); here the area measure of the elementary surfaces is sampled in closed form (see include/odw.h). */

/* point and first derivatives of the parametrisation (OCC's, see odw_face in odw.h) */
static void surface_eval(const odw_face* f, double u, double v, double* P, double* du, double* dv) {
  double cu = cos(u), su = sin(u);
  double rad[3], tan_[3];
  for (int i = 0; i < 3; ++i) { rad[i] = cu*f->xdir[i] + su*f->ydir[i]; tan_[i] = -su*f->xdir[i] + cu*f->ydir[i]; }
  switch (f->kind) {
    case ODW_SURF_PLANE:
      for (int i = 0; i < 3; ++i) { P[i] = f->origin[i] + u*f->xdir[i] + v*f->ydir[i]; du[i] = f->xdir[i]; dv[i] = f->ydir[i]; }
      break;
    case ODW_SURF_CYLINDER:
      for (int i = 0; i < 3; ++i) { P[i] = f->origin[i] + f->p0*rad[i] + v*f->zdir[i]; du[i] = f->p0*tan_[i]; dv[i] = f->zdir[i]; }
      break;
    case ODW_SURF_CONE: {
      double sa = sin(f->p1), ca = cos(f->p1), r = f->p0 + v*sa;
      for (int i = 0; i < 3; ++i) { P[i] = f->origin[i] + r*rad[i] + v*ca*f->zdir[i]; du[i] = r*tan_[i]; dv[i] = sa*rad[i] + ca*f->zdir[i]; }
      break;
    }
    case ODW_SURF_SPHERE: {
      double cv = cos(v), sv = sin(v), R = f->p0;
      for (int i = 0; i < 3; ++i) { P[i] = f->origin[i] + R*cv*rad[i] + R*sv*f->zdir[i]; du[i] = R*cv*tan_[i]; dv[i] = -R*sv*rad[i] + R*cv*f->zdir[i]; }
      break;
    }
    default: {   /* torus */
      double cv = cos(v), sv = sin(v), R = f->p0, r = f->p1;
      for (int i = 0; i < 3; ++i) { P[i] = f->origin[i] + (R + r*cv)*rad[i] + r*sv*f->zdir[i]; du[i] = (R + r*cv)*tan_[i]; dv[i] = -r*sv*rad[i] + r*cv*f->zdir[i]; }
    }
  }
}

/* (u,v) distributed by area inside the face's parameter window; returns 0 when the draw is rejected (cone, torus) */
static int surface_draw_uv(const odw_face* f, double w0, double w1, double w2, double* u, double* v) {
  double u0 = f->uv_min[0], u1 = f->uv_max[0], v0 = f->uv_min[1], v1 = f->uv_max[1];
  if (f->trim_kind == ODW_TRIM_NONE) {
    u0 = 0; u1 = TWO_PI;
    if (f->kind == ODW_SURF_SPHERE) { v0 = -TWO_PI/4; v1 = TWO_PI/4; }
    if (f->kind == ODW_SURF_TORUS) { v0 = 0; v1 = TWO_PI; }
  }
  *u = u0 + w0*(u1 - u0);
  switch (f->kind) {
    case ODW_SURF_SPHERE: {
      double s0 = sin(v0), s1 = sin(v1), sv = s0 + w1*(s1 - s0);
      if (sv > 1) sv = 1; if (sv < -1) sv = -1;
      *v = asin(sv);
      return 1;
    }
    case ODW_SURF_CONE: {
      double sa = sin(f->p1), r0 = fabs(f->p0 + v0*sa), r1 = fabs(f->p0 + v1*sa);
      double rmax = r0 > r1 ? r0 : r1;
      *v = v0 + w1*(v1 - v0);
      return w2*rmax < fabs(f->p0 + (*v)*sa);
    }
    case ODW_SURF_TORUS:
      *v = v0 + w1*(v1 - v0);
      return w2*(f->p0 + f->p1) < f->p0 + f->p1*cos(*v);
    default:
      *v = v0 + w1*(v1 - v0);
      return 1;
  }
}

/* a plane face trimmed to exactly one triangle (three chained straight pcurves; what a tessellated emitter consists of):
 * tri = (u, v) of its corners.  Such a face is sampled without rejection (include/odw.h, surface sources). */
static int emit_triangle(const odw_face* f, const odw_trimseg* segs, double* tri) {
  if (f->kind != ODW_SURF_PLANE || f->trim_kind != ODW_TRIM_LOOPS || f->seg_count != 3 || !segs) return 0;
  const odw_trimseg* g = segs + f->seg_first;
  double scale = 0;
  for (int k = 0; k < 3; ++k) { if (g[k].kind != ODW_SEG_LINE) return 0; for (int j = 0; j < 4; ++j) scale = fmax(scale, fabs(g[k].a[j])); }
  double eps = 1e-9*fmax(scale, 1e-300);
  for (int k = 0; k < 3; ++k) {
    const odw_trimseg* a = &g[k]; const odw_trimseg* b = &g[(k + 1) % 3];
    if (fabs(a->a[2] - b->a[0]) > eps || fabs(a->a[3] - b->a[1]) > eps) return 0;
    tri[2*k] = a->a[0]; tri[2*k + 1] = a->a[1];
  }
  return 1;
}

#define ODW_SURFACE_MAX_TRIES 64
static void surface_make_ray(const odw_source_desc* s, uint64_t seed, uint64_t ray, double* theta_out, double* phi_out,
                             double* origin, double* dir) {
  double a[2], b[2];
  oracle_philox(seed, (uint32_t)s->source_id, ray, 0, a);       /* a[0]: face, a[1]: theta */
  oracle_philox(seed, (uint32_t)s->source_id, ray, 1, b);       /* b[0]: phi */
  int k = 0, hi = s->n_emit-1;                                   /* first face with a[0] < emit_cdf[k] (binary search) */
  while (k < hi) { int m = (k + hi) >> 1; if (a[0] < s->emit_cdf[m]) hi = m; else k = m + 1; }
  const odw_face* f = &s->emit_faces[k];
  double P[3], du[3], dv[3], u = 0, v = 0, tri[6];
  if (emit_triangle(f, s->emit_segs, tri)) {                        /* area-uniform point of a triangle from two uniforms */
    double w[2];
    oracle_philox(seed, (uint32_t)s->source_id, ray, 2, w);
    double sq = sqrt(w[0]), b1 = sq*(1.0 - w[1]), b2 = sq*w[1], b0 = 1.0 - sq;
    u = b0*tri[0] + b1*tri[2] + b2*tri[4]; v = b0*tri[1] + b1*tri[3] + b2*tri[5];
    surface_eval(f, u, v, P, du, dv);
  } else
  for (uint32_t tr = 0; tr < ODW_SURFACE_MAX_TRIES; ++tr) {
    double w[2], w2[2];
    oracle_philox(seed, (uint32_t)s->source_id, ray, 2 + tr, w);
    oracle_philox(seed, (uint32_t)s->source_id, ray, 0x100 + tr, w2);
    int ok = surface_draw_uv(f, w[0], w[1], w2[0], &u, &v);
    surface_eval(f, u, v, P, du, dv);
    if (!ok) continue;
    double uv[2], n_tmp[3];
    surface_uv_normal(f, s->emit_segs, P, uv, n_tmp);
    if (trim_contains(f, s->emit_segs, uv, s->dist_tol)) break;   /* :399-408 keep rolling until the point is on the face */
  }
  double theta = interp_cdf(a[1], s->first_cdf, s->n_first, s->first_lo, s->first_hi);
  double phi = b[0]*TWO_PI;                                        /* :544 */
  double uv[2], n[3];
  surface_uv_normal(f, s->emit_segs, P, uv, n);
  /* :549 faceTangent = du if du.Length > 10 distTol else the longer of du, dv */
  double lu = len3(du), lv = len3(dv);
  const double* t = (lu > 10*s->dist_tol || lu >= lv) ? du : dv;
  double tl = len3(t), th[3] = { t[0]/tl, t[1]/tl, t[2]/tl }, txn[3];
  cross3(th, n, txn);
  double st = sin(theta), ct = cos(theta), sp = sin(phi), cp = cos(phi);
  double d[3];
  for (int i = 0; i < 3; ++i) d[i] = ct*n[i] + st*(cp*txn[i] + sp*th[i]);
  double dl = len3(d);
  for (int i = 0; i < 3; ++i) { origin[i] = P[i]; dir[i] = d[i]/dl; }
  if (theta_out) *theta_out = theta;
  if (phi_out) *phi_out = phi;
}

/* one Monte-Carlo ray of any source kind: the Philox stream (seed, source_id) at counter `ray` */
static void source_make_ray(const odw_source_desc* s, uint64_t seed, uint64_t ray, double* first, double* phi,
                            double* origin, double* dir) {
  if (s->kind == ODW_SRC_SURFACE) { surface_make_ray(s, seed, ray, first, phi, origin, dir); return; }
  double u[2], f, p;
  oracle_philox(seed, (uint32_t)s->source_id, ray, 0, u);
  sample_from_uniforms(s, u[0], u[1], &f, &p);
  make_ray(s, f, p, origin, dir);
  if (first) *first = f;
  if (phi) *phi = p;
}

int oracle_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(odw_face);
    case 1: return (int)sizeof(odw_trimseg);
    case 2: return (int)sizeof(odw_shell);
    case 3: return (int)sizeof(odw_group);
    case 4: return (int)sizeof(odw_scene_desc);
    case 5: return (int)sizeof(odw_source_desc);
    case 6: return (int)sizeof(odw_binning);
    case 7: return (int)sizeof(odw_trace_cfg);
    case 8: return (int)sizeof(odw_counts);
    case 9: return (int)sizeof(odw_hits_view);
  }
  return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* bounding boxes (ray.py:353-364,374,390-398): distance point->box and infinite-line/box overlap */

static double box_dist(const double* lo, const double* hi, double tol, const double* p) {
  double s = 0;
  for (int i = 0; i < 3; ++i) {
    double a = lo[i]-tol, b = hi[i]+tol;
    double d = p[i] < a ? a - p[i] : (p[i] > b ? p[i] - b : 0);
    s += d*d;
  }
  return sqrt(s);
}

static int box_line(const double* lo, const double* hi, double tol, const double* p, const double* d) {
  double t0 = -INFINITY, t1 = INFINITY;           /* FreeCAD BoundBox.intersect(base, dir): a LINE, not a half line (quirk Q4) */
  for (int i = 0; i < 3; ++i) {
    double a = lo[i]-tol, b = hi[i]+tol;
    if (d[i] == 0) { if (p[i] < a || p[i] > b) return 0; continue; }
    double ta = (a - p[i])/d[i], tb = (b - p[i])/d[i];
    if (ta > tb) { double x = ta; ta = tb; tb = x; }
    if (ta > t0) t0 = ta;
    if (tb < t1) t1 = tb;
  }
  return t0 <= t1;
}

/* ------------------------------------------------------------------------------------------ */
/* find.relevantOpticalObjects (find.py:79-104) */

static int group_relevant(const odw_scene_desc* sc, int sequential, int seq_index, int g,
                          const int32_t* ignored, int n_ignored) {
  for (int i = 0; i < n_ignored; ++i) if (ignored[i] == g) return 0;
  if (!sequential) return 1;
  if (seq_index >= sc->n_seq_steps) return 0;       /* nothing is hittable past the end of the sequence */
  for (int i = sc->seq_offsets[seq_index]; i < sc->seq_offsets[seq_index+1]; ++i)
    if (sc->seq_groups[i] == g) return 1;
  return 0;
}

typedef struct { double dist; int index; } cand_t;
static int cand_cmp(const void* a, const void* b) {
  double x = ((const cand_t*)a)->dist, y = ((const cand_t*)b)->dist;
  if (x < y) return -1; if (x > y) return 1;
  int i = ((const cand_t*)a)->index, j = ((const cand_t*)b)->index;   /* stable like Python's sorted() */
  return i < j ? -1 : (i > j);
}

typedef struct { int face; double t; double P[3]; int order; } isect_t;

/* Ray.findNearestIntersection (ray.py:290-452).  returns face index or -1 */
static int find_nearest(const odw_scene_desc* sc, const odw_trace_cfg* cfg, const double* start, const double* dir,
                        int medium, double max_len, int seq_index, const int32_t* ignored, int n_ignored,
                        cand_t* shell_c, cand_t* face_c, double* P_out, double* dist_out) {
  double tol = cfg->dist_tol;
  double dl = len3(dir);
  double dn[3] = { dir[0]/dl, dir[1]/dl, dir[2]/dl };
  isect_t hits[64]; int nh = 0;
  /* candidates: shells of relevant groups whose (enlarged) bbox is nearer than maxRayLength  (:344-364) */
  int ns = 0;
  for (int s = 0; s < sc->n_shells; ++s) {
    const odw_shell* sh = &sc->shells[s];
    if (!group_relevant(sc, cfg->sequential, seq_index, sh->group, ignored, n_ignored)) continue;
    double bd = box_dist(sh->aabb_min, sh->aabb_max, tol, start);
    if (!isfinite(max_len) || bd < max_len) { shell_c[ns].dist = bd; shell_c[ns].index = s; ++ns; }
  }
  qsort(shell_c, ns, sizeof(cand_t), cand_cmp);                                   /* :367 */
  for (int ci = 0; ci < ns; ++ci) {
    const odw_shell* sh = &sc->shells[shell_c[ci].index];
    if (!(shell_c[ci].dist < max_len && box_line(sh->aabb_min, sh->aabb_max, tol, start, dn))) continue;   /* :373-374 */
    int nf = 0;
    for (int k = 0; k < sh->face_count; ++k) {                                   /* :383-401 */
      const odw_face* f = &sc->faces[sh->face_first + k];
      double fd = box_dist(f->aabb_min, f->aabb_max, tol, start);
      if (fd < max_len && box_line(f->aabb_min, f->aabb_max, tol, start, dn)) {
        face_c[nf].dist = fd; face_c[nf].index = sh->face_first + k; ++nf;
      }
    }
    qsort(face_c, nf, sizeof(cand_t), cand_cmp);                                  /* :404 */
    for (int fi = 0; fi < nf; ++fi) {
      if (!(face_c[fi].dist < max_len)) continue;                                 /* :410 */
      const odw_face* f = &sc->faces[face_c[fi].index];
      double ts[4];
      int nt = line_surface(f, sc->segs, start, dn, ts);                                    /* :411 */
      for (int k = 0; k < nt; ++k) {
        double t = ts[k];
        double P[3] = { start[0]+t*dn[0], start[1]+t*dn[1], start[2]+t*dn[2] };
        double dist = fabs(t);                                                    /* (vec-lstart).Length */
        if (!(dist > tol)) continue;                                              /* :424 */
        double dseg = t < 0 ? -t : (t > max_len ? t - max_len : 0);               /* :425 distance to the finite segment */
        if (!(dseg < tol)) continue;
        double uv[2], nrm[3];
        surface_uv_normal(f, sc->segs, P, uv, nrm);
        if (!trim_contains(f, sc->segs, uv, tol)) continue;                       /* :426 */
        if (nh < 64) { hits[nh].face = face_c[fi].index; hits[nh].t = dist; memcpy(hits[nh].P, P, sizeof P); hits[nh].order = nh; ++nh; }
        max_len = dist + 5*tol;                                                   /* :432 */
      }
    }
  }
  if (nh == 0) return -1;                                                         /* :435 */
  double mind = hits[0].t;
  for (int i = 1; i < nh; ++i) if (hits[i].t < mind) mind = hits[i].t;
  /* keep within minDist + 2 tol, sort by distance (stable), first with group != currentMedium, else closest (:438-452) */
  int best = -1, best_other = -1;
  for (int i = 0; i < nh; ++i) {
    if (!(hits[i].t < mind + 2*tol)) continue;
    if (best < 0 || hits[i].t < hits[best].t) best = i;
    if (sc->faces[hits[i].face].group != medium)
      if (best_other < 0 || hits[i].t < hits[best_other].t) best_other = i;
  }
  int pick = best_other >= 0 ? best_other : best;
  memcpy(P_out, hits[pick].P, 3*sizeof(double));
  *dist_out = hits[pick].t;
  return hits[pick].face;
}

/* ------------------------------------------------------------------------------------------ */
/* Ray.mirror / snellsLaw / lineGrating  (ray.py:482-539) */

static void mirror(const double* ray, const double* n, double* out) {
  double k = dot3(ray, n);
  for (int i = 0; i < 3; ++i) out[i] = -(2*n[i]*k - ray[i]);
}

static int snell(const double* ray, double n1, double n2, const double* n, double* out) {
  double c[3]; cross3(n, ray, c);
  double root = 1 - n1/n2 * n1/n2 * dot3(c, c);
  if (root < 0) { mirror(ray, n, out); return 1; }
  double mn[3] = { -n[0], -n[1], -n[2] }, a[3], b[3];
  cross3(mn, ray, a); cross3(n, a, b);
  double s = sqrt(root);
  for (int i = 0; i < 3; ++i) out[i] = n1/n2*b[i] + n[i]*s;
  return 0;
}

/* unpolarised Fresnel reflectance of the interface n1 -> n2 for unit direction d and unit normal n (d.n >= 0); 1 beyond the
 * critical angle.  Not in the reference (no Fresnel split there): opt-in extension, see odw_group.fresnel in odw.h */
static double fresnel_reflectance(const double* d, const double* n, double n1, double n2) {
  double ci = fabs(dot3(d, n));
  if (ci > 1) ci = 1;
  double s2 = (n1/n2)*(n1/n2)*(1 - ci*ci);
  if (s2 >= 1) return 1.0;
  double ct = sqrt(1 - s2);
  double rs = (n1*ci - n2*ct)/(n1*ci + n2*ct), rp = (n1*ct - n2*ci)/(n1*ct + n2*ci);
  return 0.5*(rs*rs + rp*rp);
}

static void line_grating(const double* ray_in, double n1, double n2, const double* normal, const odw_group* g,
                         double wavelength_nm, int transmission, double* out) {
  double wl = wavelength_nm/1000.0;
  double rl = len3(ray_in), nl = len3(normal), gl = len3(g->grating_orientation);
  double ray[3], sn[3], gv[3];
  for (int i = 0; i < 3; ++i) { ray[i] = ray_in[i]/rl; sn[i] = normal[i]/nl; gv[i] = g->grating_orientation[i]/gl; }
  double P[3], D[3];
  cross3(gv, sn, P); double pl = len3(P); for (int i = 0; i < 3; ++i) P[i] /= pl;
  cross3(sn, P, D);  double dl = len3(D); for (int i = 0; i < 3; ++i) D[i] /= dl;
  double mu = n1/n2, d = 1000.0/g->grating_lines_per_mm;
  double T = (g->grating_order*wl)/(n1*d);
  double nn = dot3(sn, sn);
  double V = (mu*dot3(ray, sn))/nn;
  double W = (mu*mu - 1 + T*T - 2*mu*T*dot3(ray, D))/nn;
  double sq = sqrt((2*V)*(2*V) - 4*W);                 /* NaN propagates like the reference's complex/NaN result */
  double q1 = (-2*V + sq)/2, q2 = (-2*V - sq)/2;
  double Q = transmission ? fmin(q1, q2) : fmax(q1, q2);
  for (int i = 0; i < 3; ++i) out[i] = -(mu*ray[i] - T*D[i] + Q*sn[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* hit sink */

typedef struct {
  odw_hits_view* view;
  uint64_t* n_hits;       /* shared counter */
  uint64_t* dropped;
  double* bins;           /* concatenated histograms */
  const odw_trace_cfg* cfg;
} sink_t;

static void bin_hit(const sink_t* sk, int group, const double* P, double power) {
  size_t off = 0;
  for (int b = 0; b < sk->cfg->n_binnings; ++b) {
    const odw_binning* bn = &sk->cfg->binnings[b];
    size_t nb = (size_t)bn->nu*(size_t)bn->nv;
    if (bn->group == group && sk->bins) {
      double w[3] = { P[0]-bn->origin[0], P[1]-bn->origin[1], P[2]-bn->origin[2] };
      double x = dot3(w, bn->uaxis), y = dot3(w, bn->vaxis);
      if (x >= bn->u_lo && x <= bn->u_hi && y >= bn->v_lo && y <= bn->v_hi) {
        int ix = (int)((x - bn->u_lo)*bn->nu/(bn->u_hi - bn->u_lo)); if (ix >= bn->nu) ix = bn->nu-1;
        int iy = (int)((y - bn->v_lo)*bn->nv/(bn->v_hi - bn->v_lo)); if (iy >= bn->nv) iy = bn->nv-1;
        double add = bn->weighted ? power : 1.0;
#ifdef _OPENMP
#pragma omp atomic
#endif
        sk->bins[off + (size_t)ix*bn->nv + iy] += add;
      }
    }
    off += nb;
  }
}

static void record_hit(const sink_t* sk, uint64_t ray_index, int bounce, int group, int face_id,
                       const double* P, const double* dir, double power, int entering, int medium) {
  bin_hit(sk, group, P, power);
  uint64_t slot;
#ifdef _OPENMP
#pragma omp atomic capture
#endif
  slot = (*sk->n_hits)++;
  odw_hits_view* v = sk->view;
  if (!sk->cfg->store_hits || !v) return;
  if (slot >= v->capacity) {
#ifdef _OPENMP
#pragma omp atomic
#endif
    (*sk->dropped)++;
    return;
  }
  if (v->points)      memcpy(v->points + 3*slot, P, 3*sizeof(double));
  if (v->directions)  memcpy(v->directions + 3*slot, dir, 3*sizeof(double));
  if (v->powers)      v->powers[slot] = power;
  if (v->is_entering) v->is_entering[slot] = (uint8_t)entering;
  if (v->ray_index)   v->ray_index[slot] = ray_index;
  if (v->group)       v->group[slot] = group;
  if (v->bounce)      v->bounce[slot] = bounce;
  if (v->face_id)     v->face_id[slot] = face_id;
  if (v->medium)      v->medium[slot] = medium;
}

/* ------------------------------------------------------------------------------------------ */
/* Ray.traceRay (ray.py:36-281) */

typedef struct { uint64_t segments, escaped, depth_terminated; } ray_stats_t;

static void trace_one(const odw_scene_desc* sc, const odw_trace_cfg* cfg, const double* origin, const double* dir0,
                      double power0, double wavelength, double max_len, int max_isect,
                      const int32_t* ignored, int n_ignored, uint64_t ray_index, uint64_t seed, uint32_t source_id, const sink_t* sk,
                      cand_t* shell_c, cand_t* face_c, ray_stats_t* st,
                      int32_t* n_segments, double* final_point, double* final_power, int32_t* final_medium) {
  double point[3] = { origin[0], origin[1], origin[2] };
  double dir[3] = { dir0[0], dir0[1], dir0[2] };
  double power = power0;
  int medium = -1;                 /* currentMedium: None */
  int seq_index = 0, n_isect = 0, nseg = 0;
  for (;;) {
    if (n_isect >= max_isect) { st->depth_terminated++; break; }            /* :96-98 */
    n_isect++;
    double P[3], dist;
    int fi = find_nearest(sc, cfg, point, dir, medium, max_len, seq_index, ignored, n_ignored,
                          shell_c, face_c, P, &dist);                       /* :101 */
    if (fi < 0) {                                                           /* :105-109 final segment */
      double dl = len3(dir);
      for (int i = 0; i < 3; ++i) point[i] += dir[i]/dl*max_len;
      nseg++; st->escaped++;
      break;
    }
    const odw_face* f = &sc->faces[fi];
    const odw_group* g = &sc->groups[f->group];
    double prev[3] = { point[0], point[1], point[2] };
    int prev_medium = medium;
    memcpy(point, P, sizeof P);
    nseg++;                                                                 /* :117 yield */
    if (prev_medium >= 0) {                                                 /* :120-125 absorption (quirk Q1: multiplicative here) */
      double L = sc->groups[prev_medium].absorption_length;
      if (L == 0) power = 0;
      else if (isfinite(L)) {
        double dd[3] = { prev[0]-point[0], prev[1]-point[1], prev[2]-point[2] };
        power *= exp(-len3(dd)/L);
      }
    }
    /* getNormal (:455-480): outward normal, flipped to point along propagation; isEntering */
    double uv[2], n_out[3], nrm[3];
    surface_uv_normal(f, sc->segs, point, uv, n_out);
    double dray[3] = { point[0]-prev[0], point[1]-prev[1], point[2]-prev[2] };
    double cosang = dot3(dray, n_out)/(len3(dray)*len3(n_out));
    int entering = cosang < 0;
    for (int i = 0; i < 3; ++i) nrm[i] = entering ? -n_out[i] : n_out[i];
    /* onRayHit (:131; optical_group.py:206-209) */
    if (g->record_hits || cfg->record_all_hits)
      record_hit(sk, ray_index, n_isect-1, f->group, f->face_id, point, dir, power, entering, prev_medium);
    double dl = len3(dir);
    double dnrm[3] = { dir[0]/dl, dir[1]/dl, dir[2]/dl };
    switch (g->optical_type) {
      case ODW_OPT_MIRROR: {                                                /* :146-161 */
        double o[3]; mirror(dir, nrm, o);
        apply_scatter(sc, f->group, seed, source_id, ray_index, n_isect-1, dnrm, nrm, entering, o);   /* :151-155 */
        memcpy(dir, o, sizeof o);
        power *= g->reflectivity;
        seq_index++;
        break;
      }
      case ODW_OPT_LENS: {                                                  /* :165-211 */
        double n1, n2;
        if (entering) {
          n1 = medium >= 0 ? sc->groups[medium].refractive_index : 1.0;
          medium = f->group;
          n2 = sc->groups[medium].refractive_index;
        } else {
          n1 = medium >= 0 ? sc->groups[medium].refractive_index : 1.0;
          n2 = 1.0;                                                         /* quirk Q2 */
        }
        double o[3];
        if (g->fresnel) {                                                   /* opt-in extension (odw.h odw_group.fresnel): not in the reference */
          double u[2];
          oracle_philox(seed, source_id, ray_index, 0x20000u + (uint32_t)(n_isect-1), u);
          if (u[0] < fresnel_reflectance(dnrm, nrm, n1, n2)) {
            if (entering) medium = prev_medium;                             /* the ray never got in */
            mirror(dnrm, nrm, o);
            memcpy(dir, o, sizeof o);
            break;
          }
        }
        int tir = snell(dnrm, n1, n2, nrm, o);
        apply_scatter(sc, f->group, seed, source_id, ray_index, n_isect-1, dnrm, nrm, entering, o);   /* :197-201 */
        memcpy(dir, o, sizeof o);
        if (!entering && !tir && medium == f->group) { medium = -1; seq_index++; }
        break;
      }
      case ODW_OPT_GRATING: {                                               /* :216-268 */
        if (g->grating_type == ODW_GRATING_REFLECTION) {
          if (entering) {
            double n = medium >= 0 ? sc->groups[medium].refractive_index : 1.0;
            double o[3]; line_grating(dnrm, n, n, nrm, g, wavelength, 0, o); memcpy(dir, o, sizeof o);
            seq_index++;
          }
        } else {
          if (entering) {
            if (medium >= 0) { power = 0; break; }                          /* reference raises ValueError here */
            medium = f->group;
            double o[3]; line_grating(dnrm, 1.0, g->refractive_index, nrm, g, wavelength, 1, o); memcpy(dir, o, sizeof o);
          } else {
            double n1 = medium >= 0 ? sc->groups[medium].refractive_index : 1.0;
            double o[3];
            int tir = snell(dnrm, n1, 1.0, nrm, o);
            memcpy(dir, o, sizeof o);
            if (!tir) { medium = -1; seq_index++; }
          }
        }
        break;
      }
      case ODW_OPT_ABSORBER: power = 0; seq_index++; break;                 /* :271-273 */
      case ODW_OPT_VACUUM: seq_index++; break;                              /* :276-277 */
    }
    if (power < cfg->power_tol) break;                                      /* :280 */
  }
  st->segments += nseg;
  if (n_segments) *n_segments = nseg;
  if (final_point) memcpy(final_point, point, 3*sizeof(double));
  if (final_power) *final_power = power;
  if (final_medium) *final_medium = medium;
}

/* ------------------------------------------------------------------------------------------ */
/* drivers */

static size_t total_bins(const odw_trace_cfg* cfg) {
  size_t n = 0;
  for (int b = 0; b < cfg->n_binnings; ++b) n += (size_t)cfg->binnings[b].nu*(size_t)cfg->binnings[b].nv;
  return n;
}

int oracle_trace_rays(const odw_scene_desc* sc, const odw_trace_cfg* cfg,
                      const double* origins, const double* dirs, const double* powers, double wavelength,
                      const int32_t* ignored, int32_t n_ignored, uint64_t n, uint64_t ray_index_base,
                      odw_hits_view* hits, uint64_t* n_hits_out, odw_counts* counts,
                      int32_t* n_segments, double* final_points, double* final_powers, int32_t* final_media, double* bins,
                      int n_threads) {
  uint64_t n_hits = 0, dropped = 0;
  uint64_t segs = 0, esc = 0, depth = 0;
  if (bins) memset(bins, 0, total_bins(cfg)*sizeof(double));
  sink_t sk = { hits, &n_hits, &dropped, bins, cfg };
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel reduction(+:segs,esc,depth)
#endif
  {
    cand_t* shell_c = (cand_t*)malloc(sizeof(cand_t)*(size_t)(sc->n_shells + 1));
    cand_t* face_c  = (cand_t*)malloc(sizeof(cand_t)*(size_t)(sc->n_faces + 1));
    ray_stats_t st = {0, 0, 0};
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      trace_one(sc, cfg, origins + 3*i, dirs + 3*i, powers ? powers[i] : 1.0, wavelength,
                cfg->max_ray_length, cfg->max_intersections, ignored, n_ignored, ray_index_base + (uint64_t)i,
                cfg->scatter_seed, 0u, &sk,
                shell_c, face_c, &st,
                n_segments ? n_segments + i : NULL, final_points ? final_points + 3*i : NULL,
                final_powers ? final_powers + i : NULL, final_media ? final_media + i : NULL);
    }
    segs += st.segments; esc += st.escaped; depth += st.depth_terminated;
    free(shell_c); free(face_c);
  }
  if (n_hits_out) *n_hits_out = n_hits - dropped;
  if (counts) {
    memset(counts, 0, sizeof *counts);
    counts->rays = n; counts->segments = segs; counts->hits = n_hits; counts->hits_dropped = dropped;
    counts->escaped = esc; counts->depth_terminated = depth; counts->waves = 1;
  }
  return dropped ? ODW_EOVERFLOW : 0;
}

int oracle_trace_mc(const odw_scene_desc* sc, const odw_source_desc* src, const odw_trace_cfg* cfg,
                    uint64_t seed, uint64_t first_ray, uint64_t n,
                    odw_hits_view* hits, uint64_t* n_hits_out, odw_counts* counts, double* bins, int n_threads) {
  uint64_t n_hits = 0, dropped = 0;
  uint64_t segs = 0, esc = 0, depth = 0;
  if (bins) memset(bins, 0, total_bins(cfg)*sizeof(double));
  sink_t sk = { hits, &n_hits, &dropped, bins, cfg };
  double max_len = cfg->max_ray_length*src->max_ray_length_scale;            /* ray.py:48-53 */
  int max_isect = (int)(cfg->max_intersections*src->max_intersections_scale);
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel reduction(+:segs,esc,depth)
#endif
  {
    cand_t* shell_c = (cand_t*)malloc(sizeof(cand_t)*(size_t)(sc->n_shells + 1));
    cand_t* face_c  = (cand_t*)malloc(sizeof(cand_t)*(size_t)(sc->n_faces + 1));
    ray_stats_t st = {0, 0, 0};
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      double o[3], d[3];
      uint64_t ray = first_ray + (uint64_t)i;
      source_make_ray(src, seed, ray, NULL, NULL, o, d);
      trace_one(sc, cfg, o, d, 1.0, src->wavelength, max_len, max_isect, src->ignored_groups, src->n_ignored,
                ray, seed, (uint32_t)src->source_id, &sk, shell_c, face_c, &st, NULL, NULL, NULL, NULL);
    }
    segs += st.segments; esc += st.escaped; depth += st.depth_terminated;
    free(shell_c); free(face_c);
  }
  if (n_hits_out) *n_hits_out = n_hits - dropped;
  if (counts) {
    memset(counts, 0, sizeof *counts);
    counts->rays = n; counts->segments = segs; counts->hits = n_hits; counts->hits_dropped = dropped;
    counts->escaped = esc; counts->depth_terminated = depth; counts->waves = 1;
  }
  return dropped ? ODW_EOVERFLOW : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* The two geometry questions the reference asks OpenCASCADE, exported one at a time: tests/golden/make_traceray_golden.py
 * runs the reference's OWN Ray.traceRay (ray.py:36-281) with these as its geometry provider. */

/* Ray.findNearestIntersection (ray.py:290-452) for one segment; returns the face index or -1 */
int oracle_find_nearest(const odw_scene_desc* sc, const odw_trace_cfg* cfg, const double* start, const double* dir,
                        int32_t medium, double max_len, int32_t seq_index, const int32_t* ignored, int32_t n_ignored,
                        double* point_out) {
  cand_t* shell_c = (cand_t*)malloc(sizeof(cand_t)*(size_t)(sc->n_shells + 1));
  cand_t* face_c  = (cand_t*)malloc(sizeof(cand_t)*(size_t)(sc->n_faces + 1));
  double dist;
  int fi = find_nearest(sc, cfg, start, dir, medium, max_len, seq_index, ignored, n_ignored, shell_c, face_c, point_out, &dist);
  free(shell_c); free(face_c);
  return fi;
}

/* Surface.parameter + Face.normalAt (ray.py:463-466): (u, v) and the orientation-aware unit normal of a face at P */
int oracle_face_normal(const odw_scene_desc* sc, int32_t face, const double* P, double* uv_out, double* normal_out) {
  if (face < 0 || face >= sc->n_faces) return ODW_EINVAL;
  surface_uv_normal(&sc->faces[face], sc->segs, P, uv_out, normal_out);
  return 0;
}

/* the (theta, phi) applyStochasticRayCorrections draws for interaction `bounce` of ray `ray` on `group`
 * (which = 0: Reflected/RefractedProbabilityDensity, 1: RayModificationProbabilityDensity); 0 = that density is empty */
int oracle_scatter_draw(const odw_scene_desc* sc, int32_t group, int32_t which, uint64_t seed, uint32_t source_id,
                        uint64_t ray, int32_t bounce, double* theta, double* phi) {
  if (!sc->n_scatters || !sc->group_scatter || group < 0 || group >= sc->n_groups) return 0;
  int idx = sc->group_scatter[2*group + (which ? 1 : 0)];
  if (idx < 0) return 0;
  double u[2];
  oracle_philox(seed, source_id, ray, 0x10000u + 4u*(uint32_t)bounce + (which ? 1u : 0u), u);
  scatter_sample(&sc->scatters[idx], u[0], u[1], theta, phi);
  return 1;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
