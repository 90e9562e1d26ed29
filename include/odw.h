/*
 * odw.h — C ABI of the B200-native ray-trace engine for the FreeCAD Optics Design Workbench.
 *
 * This header is the drop-in boundary for ONE hot path of zaphB/freecad.optics_design_workbench:
 * the per-simulation trace loop.  All citations are relative to the reference tree
 * (freecad/optics_design_workbench/…):
 *
 *   freecad_elements/generic_source.py:51-146   GenericSourceProxy.runSimulationIteration  (the call site replaced)
 *   freecad_elements/ray.py:36-281              Ray.traceRay            (bounce loop)       -> odw_trace_mc / odw_trace_rays
 *   freecad_elements/ray.py:290-452             Ray.findNearestIntersection                 -> same kernels
 *   freecad_elements/ray.py:455-539             getNormal / mirror / snellsLaw / lineGrating -> same kernels
 *   freecad_elements/point_source.py:411-460    PointSourceProxy._makeRay                   -> odw_source_create + odw_trace_mc
 *   distributions/random_number_generator.py:372-560  tabulated inverse-CDF sampler         -> odw_source_create (tables) + odw_trace_mc (draw)
 *   freecad_elements/optical_group.py:206-209   OpticalGroupProxy.onRayHit                  -> hit append (odw_result_hits)
 *   simulation/results_store.py:641-648         SimulationResults.addRayHit                 -> hit append (odw_result_hits)
 *   jupyter_utils/hits.py:176-193, histogram.py:24-85  Hits.histogram (post-hoc in reference) -> odw_result_histogram (device binning)
 *
 * Conventions: every function returns 0 on success and a negative ODW_E* code on failure;
 * odw_last_error() returns a thread-local message.  No exceptions cross the boundary.  Handles
 * are opaque and freed by the matching *_destroy.  All arrays are caller-allocated, contiguous,
 * host memory; floating point is IEEE fp64 everywhere (the reference computes in Python floats /
 * FreeCAD Vector doubles).  Units: mm, nm (wavelength), rad.  A handle may be used from one
 * thread at a time.
 *
 * The library has NO CPU fallback: odw_engine_create fails (ODW_ENODEVICE) without a CUDA device.
 * The CPU restatement used by the tests lives in oracle/ and is not linked here.
 */
#ifndef ODW_H
#define ODW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODW_ABI_VERSION 2

/* error codes */
#define ODW_OK           0
#define ODW_EINVAL      -1   /* bad argument / malformed description */
#define ODW_ENODEVICE   -2   /* no CUDA device / device id out of range */
#define ODW_ECUDA       -3   /* CUDA runtime error (message has the detail) */
#define ODW_ENOMEM      -4   /* host or device allocation failed */
#define ODW_EOVERFLOW   -5   /* hit buffer too small: result is complete except dropped hits (see odw_counts.hits_dropped) */
#define ODW_EUNSUPPORTED -6

/* surface kinds (numbering follows the OCC BRep surface type ids 1..5) */
#define ODW_SURF_PLANE     1
#define ODW_SURF_CYLINDER  2
#define ODW_SURF_CONE      3
#define ODW_SURF_SPHERE    4
#define ODW_SURF_TORUS     5
#define ODW_SURF_CONICOID  6   /* conic of revolution (paraboloid, ellipsoid, hyperboloid sheet, sphere cap), see odw_face */

/* trim kinds: how "point lies on the trimmed face" (ray.py:426) is decided in the surface's (u,v) space */
#define ODW_TRIM_NONE   0    /* whole (closed) surface, e.g. full sphere / full torus */
#define ODW_TRIM_UVBOX  1    /* uv_min <= (u,v) <= uv_max, u (and v for a torus) reduced into the period window starting at uv_min */
#define ODW_TRIM_LOOPS  2    /* even-odd rule over the face's boundary pcurves (odw_trimseg list) */

/* trim segment kinds (pcurves of the face boundary in (u,v) space) */
#define ODW_SEG_LINE 1       /* a = {u0, v0, u1, v1} */
#define ODW_SEG_ARC  2       /* a = {cu, cv, radius, angle0, span}: angles in [angle0, angle0+span], span in (0, 2pi] */
#define ODW_SEG_ASPHERE 3    /* not a boundary piece: auxiliary record of a conicoid face, a = even-asphere coefficients of
                              * rho^4, rho^6, rho^8, rho^10, rho^12 added to the conic sag.  Must be the face's FIRST segment
                              * (seg_first; counted in seg_count, also when trim_kind is not ODW_TRIM_LOOPS) */

/* optical types, order of the reference's OpticalType enumeration (optical_group.py:29-96) */
#define ODW_OPT_MIRROR   0
#define ODW_OPT_LENS     1
#define ODW_OPT_GRATING  2
#define ODW_OPT_ABSORBER 3
#define ODW_OPT_VACUUM   4

#define ODW_GRATING_REFLECTION   0
#define ODW_GRATING_TRANSMISSION 1

/* source kinds */
#define ODW_SRC_POINT_SPHERICAL  0   /* finite focal length: (theta, phi) sampling, point_source.py:424-435 */
#define ODW_SRC_POINT_COLLIMATED 1   /* FocalLength = inf:   (r, phi) sampling,     point_source.py:438-446 */
#define ODW_SRC_SURFACE          2   /* SurfaceSourceProxy: emission from faces,    surface_source.py:418-555 */

/* One face instance in WORLD coordinates (placements/links/arrays already applied by the scene export:
 * M = gpM * pMi of ray.py:338-339 folded into origin/xdir/ydir/zdir).
 * Parametrisations are OCC's:  plane O+u X+v Y;  cylinder O+r(cos u X+sin u Y)+v Z;
 * cone O+(r+v sin a)(cos u X+sin u Y)+v cos a Z;  sphere O+R cos v(cos u X+sin u Y)+R sin v Z;
 * torus O+(R+r cos v)(cos u X+sin u Y)+r sin v Z;
 * conicoid O+v(cos u X+sin u Y)+sag(v) Z with the optical sag sag(v) = c v^2/(1+sqrt(1-(1+k) c^2 v^2)), v = distance from
 * the axis >= 0, c = vertex curvature (p0), k = conic constant (p1): k = -1 paraboloid of focal length 1/(2c) (what OCC
 * writes as the surface of revolution of a Geom_Parabola about its own axis), -1 < k < 0 prolate / k > 0 oblate
 * ellipsoid, k < -1 hyperboloid sheet, k = 0 sphere.  n_geom of a conicoid = du x dv / |du x dv| ~ c rho_vec - q Z,
 * q = 1 - (1+k) c z (the side the vertex bulges towards for c > 0).  With an ODW_SEG_ASPHERE record the sag gains
 * a4 v^4 + ... + a12 v^12 (standard even asphere); the ray crossing is then found by Newton from the crossings of the base conic.
 * Outward normal n_out = nsign * n_geom, n_geom = the radial-outward normal written with (X,Y,Z)
 * (plane: Z); nsign folds the face orientation (TopAbs_REVERSED) and the handedness of the frame. */
typedef struct odw_face {
  double origin[3];
  double xdir[3];
  double ydir[3];
  double zdir[3];
  double p0;            /* cylinder r | cone r (at v=0) | sphere R | torus major R | conicoid vertex curvature c */
  double p1;            /* cone semi-angle | torus minor r | conicoid conic constant k */
  double uv_min[2];     /* trim bounding box in (u,v); also the start of the period window */
  double uv_max[2];
  double aabb_min[3];   /* world AABB of the trimmed face (not enlarged; tolerance added by the consumer) */
  double aabb_max[3];
  int32_t kind;         /* ODW_SURF_* */
  int32_t trim_kind;    /* ODW_TRIM_* */
  int32_t nsign;        /* +1 / -1 */
  int32_t group;        /* index into odw_scene_desc.groups */
  int32_t shell;        /* index into odw_scene_desc.shells */
  int32_t seg_first;    /* first trim segment (ODW_TRIM_LOOPS) */
  int32_t seg_count;
  int32_t face_id;      /* exporter's id of the source face (diagnostics / parity reports) */
} odw_face;

typedef struct odw_trimseg {
  double a[5];
  int32_t kind;         /* ODW_SEG_* */
  int32_t pad;
} odw_trimseg;

/* One shell instance (ray.py:345-364 culls per shell bounding box first). */
typedef struct odw_shell {
  double aabb_min[3];
  double aabb_max[3];
  int32_t face_first;
  int32_t face_count;
  int32_t group;
  int32_t pad;
} odw_shell;

/* One optical group = one OpticalGroup document object (optical_group.py:29-96). */
typedef struct odw_group {
  double refractive_index;
  double reflectivity;
  double absorption_length;      /* +inf = transparent (AbsorptionLength 'inf') */
  double grating_lines_per_mm;
  double grating_order;
  double grating_orientation[3]; /* GratingLinesOrientation, world frame */
  int32_t optical_type;          /* ODW_OPT_* */
  int32_t record_hits;           /* RecordHits */
  int32_t grating_type;          /* ODW_GRATING_* */
  int32_t fresnel;               /* OPT-IN EXTENSION, 0 = the reference's behaviour (ray.py:165-211 refracts every ray at a Lens
                                  * face without loss: no Fresnel split).  1: at every face of this Lens group the
                                  * ray is REFLECTED with the unpolarised Fresnel reflectance R = (Rs + Rp)/2 of the interface
                                  * n1 -> n2 and refracted otherwise (one ray in, one ray out: a stochastic choice driven by the
                                  * ray's Philox stream, purpose 0x20000 + bounce; power is not split).  A reflected ray keeps its
                                  * medium and its sequence index, like a totally reflected one. */
} odw_group;

/* Tabulated (theta, phi) density of an optical group's stochastic surface model (optical_group.py:212-323:
 * ReflectedProbabilityDensity of a mirror, RefractedProbabilityDensity of a lens, RayModificationProbabilityDensity of
 * either).  Same table layout and draw as a point source sampler (phi from the marginal, theta from the row of the
 * nearest phi mid-point); NO sin(theta) factor is added (optical_group.py:219-223).
 * n_tables > 1 (main density of a group only): the density depends on the incidence of the hit (theta_in, theta_refl of
 * optical_group.py:288-307, which the reference substitutes for every hit) and is given as a FAMILY of tables over
 * theta_in = angle(incoming direction, normal along the propagation) on linspace(0, pi/2, K): K = n_tables for a Mirror
 * group, n_tables / 2 for a Lens group (first K tables: entering hits, then K for leaving hits; theta_refl follows from
 * theta_in by the law of reflection / refraction).  A hit draws from the member nearest to its theta_in.  n_rows must be 1. */
typedef struct odw_scatter {
  int32_t n_first, n_phi, n_rows;
  int32_t n_tables;              /* 0 or 1: one table */
  double first_lo, first_hi, phi_lo, phi_hi;
  const double* phi_cdf;         /* [n_tables][n_phi] */
  const double* first_cdf;       /* [n_tables][n_rows][n_first] */
} odw_scatter;

typedef struct odw_scene_desc {
  int32_t n_faces;
  int32_t n_segs;
  int32_t n_shells;
  int32_t n_groups;
  int32_t n_seq_steps;           /* number of non-empty SequentialModeElements_NN lists (simulation_settings.py:158-196) */
  int32_t n_seq_entries;
  const odw_face*    faces;      /* sorted by shell */
  const odw_trimseg* segs;
  const odw_shell*   shells;
  const odw_group*   groups;
  const int32_t*     seq_offsets; /* [n_seq_steps+1] into seq_groups */
  const int32_t*     seq_groups;  /* group indices */
  /* stochastic surface models (applyStochasticRayCorrections, optical_group.py:279-323); n_scatters = 0: all ideal.
   * After the ideal mirror / Snell direction of a Mirror or Lens hit:
   *   main density   (theta, phi) -> d = cos(theta) n + sin(theta) (cos(phi) (a x n) + sin(phi) a),  a = unit(n x d_in)
   *                  [= Rotation(n, phi) Rotation(n x d_in, theta) n, n = face normal flipped along the propagation]
   *   modify density (theta, phi) -> the same formula with n replaced by the current outgoing direction
   * Draws: Philox stream of the ray, purposes 0x10000 + 4*bounce (main) and + 1 (modify). */
  int32_t n_scatters;
  int32_t pad0;
  const odw_scatter* scatters;
  const int32_t*     group_scatter; /* [n_groups][2] = {main, modify} index into scatters, -1 = ideal; NULL = all ideal */
} odw_scene_desc;

/* Point source + its tabulated sampler (random_number_generator.py:372-464, Appendix D of SURVEY.md).
 * Edges are linspace(lo, hi, n) like the reference.  phi is drawn first from the marginal CDF, then
 * theta (or r) from the row of the conditional table whose phi mid-point is nearest (argmin |C_phi - phi|).
 * n_rows == 1 declares the density phi-independent (one shared conditional row). */
typedef struct odw_source_desc {
  int32_t kind;                  /* ODW_SRC_* */
  int32_t source_id;             /* index of the light source; part of the Philox key */
  int32_t n_first;               /* number of edges of the first variable (theta | r) */
  int32_t n_phi;                 /* number of phi edges */
  int32_t n_rows;                /* n_phi-1, or 1 */
  int32_t n_ignored;             /* IgnoredOpticalElements (generic_source.py:23-37) */
  double first_lo, first_hi;     /* theta (or r) domain */
  double phi_lo, phi_hi;
  double focal_length;           /* ignored for the collimated kind */
  double wavelength;             /* nm */
  double max_ray_length_scale;   /* MaxRayLengthScale */
  double max_intersections_scale;/* MaxIntersectionsScale */
  double gpM[16];                /* row-major 4x4 global placement of the source */
  const double* phi_cdf;         /* [n_phi]  normalised to cdf[n_phi-1] == 1 */
  const double* first_cdf;       /* [n_rows][n_first] each row normalised */
  const int32_t* ignored_groups; /* [n_ignored] */
  /* ---- surface sources only (kind == ODW_SRC_SURFACE; reference freecad_elements/surface_source.py:418-555) ----
   * Per ray: pick an emitting face with probability proportional to its area (:465-466,536-537), draw a point
   * uniformly by area inside the face's (u,v) window and redraw until it lies on the trimmed face within dist_tol
   * (:390-410; a plane face trimmed to a single triangle — the tessellation of a free-form emitter — is sampled directly,
   * P = (1 - sqrt(w0)) A + sqrt(w0) (1 - w1) B + sqrt(w0) w1 C, nothing to redraw), draw theta from first_cdf (one row, edges linspace(first_lo, first_hi, n_first); NO sin(theta)
   * factor, :530) and phi uniform in [0, 2pi) (:544), then
   *   d = cos(theta) n + sin(theta) (cos(phi) (t x n) + sin(phi) t)        [= R(n,phi) R(t,theta) n, :85-111]
   * with n the outward face normal and t the unit u-tangent (the longer of the u/v tangents when |dP/du| <= 10 dist_tol).
   * phi_cdf, n_phi, n_rows, focal_length and gpM are ignored (the faces are given in WORLD coordinates: gpM*pMi of the
   * emitting part already applied).  The reference tabulates the area element on an adaptively refined (u,v) grid
   * (:269-387); the engine samples the area measure of the elementary surfaces in closed form instead — the same
   * distribution without the table's discretisation error. */
  int32_t n_emit;                /* emitting face instances */
  int32_t n_emit_segs;
  const odw_face* emit_faces;    /* [n_emit], world frame; group/shell fields unused */
  const odw_trimseg* emit_segs;  /* [n_emit_segs] trim loops of the emitting faces */
  const double* emit_cdf;        /* [n_emit] cumulative area weights, emit_cdf[n_emit-1] == 1 */
  double dist_tol;               /* on-face tolerance, max(DistanceTolerance, 1e-9) (:113-119) */
} odw_source_desc;

/* Detector binning (new capability; semantics = numpy.histogram2d over plane-projected hit points as in
 * jupyter_utils/histogram.py:54, but with an explicit plane instead of hits.py:96-174 auto-detection). */
typedef struct odw_binning {
  int32_t group;                 /* hits of this optical group are binned */
  int32_t nu, nv;
  int32_t weighted;              /* 0: counts, 1: sum of hit power */
  double origin[3];
  double uaxis[3];               /* x = (P-origin).uaxis */
  double vaxis[3];
  double u_lo, u_hi, v_lo, v_hi; /* numpy.histogram2d range; last bin closed on the right */
} odw_binning;

typedef struct odw_trace_cfg {
  double max_ray_length;         /* MaxRayLength (settings), multiplied by the source's scale for MC */
  double dist_tol;               /* max(DistanceTolerance, 1e-6), ray.py:283-288 */
  double power_tol;              /* 1e-6, ray.py:36 */
  int32_t max_intersections;     /* MaxIntersections */
  int32_t sequential;            /* SequentialMode */
  int32_t record_all_hits;       /* 1: record every intersection regardless of RecordHits (parity runs) */
  int32_t store_hits;            /* 0: count only (+ binning), 1: keep hit lists */
  int32_t n_binnings;
  int32_t bounces_per_wave;      /* 0 = engine default; rays alive after this many bounces are compacted into the next wave */
  uint64_t hit_capacity;         /* 0 = engine default (n_rays * 2) */
  const odw_binning* binnings;
  double wavelength;             /* odw_trace_rays only: wavelength (nm) of the listed rays (gratings); 0 = 500.  odw_trace_mc uses the source's */
  uint64_t scatter_seed;         /* odw_trace_rays only: Philox key of the stochastic-surface draws (ray = row of the list,
                                    source id 0); odw_trace_mc uses its seed argument and the source's id */
} odw_trace_cfg;

typedef struct odw_counts {
  uint64_t rays;                 /* totalTracedRays   (generic_source.py:141) */
  uint64_t segments;             /* traceRay yields   (ray.py:107,117) */
  uint64_t hits;                 /* recorded hits     (results_store.py:641-648) */
  uint64_t hits_dropped;         /* hits that did not fit hit_capacity */
  uint64_t escaped;              /* rays ending with a no-intersection segment (ray.py:105-109) */
  uint64_t depth_terminated;     /* rays stopped by maxIntersections (ray.py:96-98) */
  uint64_t waves;                /* kernel launches of the bounce loop */
  uint64_t sm_clock_khz;         /* effective SM clock observed inside the kernel (clock64 / globaltimer of CTA 0), kHz */
} odw_counts;

/* Host copy-out target for hit lists.  Any pointer may be NULL (that column is skipped). */
typedef struct odw_hits_view {
  uint64_t capacity;             /* rows available in each non-NULL array */
  double*   points;              /* [n][3] */
  double*   directions;          /* [n][3] incoming direction (ray.py:131-134) */
  double*   powers;              /* [n] */
  uint8_t*  is_entering;         /* [n] */
  uint64_t* ray_index;           /* [n] global ray index (MC: Philox counter; explicit list: row) */
  int32_t*  group;               /* [n] optical group index */
  int32_t*  bounce;              /* [n] 0-based intersection number along the ray */
  int32_t*  face_id;             /* [n] */
  int32_t*  medium;              /* [n] optical group the segment ENDING at this hit travelled through, -1 = none (currentMedium of
                                    ray.py:117 at the yield; feeds the `media` list of *-rays.pkl, results_store.py:241-257) */
} odw_hits_view;

typedef struct odw_engine odw_engine;
typedef struct odw_scene  odw_scene;
typedef struct odw_source odw_source;
typedef struct odw_result odw_result;

int  odw_abi_version(void);
const char* odw_last_error(void);

int  odw_engine_create(int device_id, odw_engine** out);
void odw_engine_destroy(odw_engine*);
int  odw_engine_device_name(const odw_engine*, char* buf, int buflen);
/* the cudaStream_t every kernel/copy of this engine is issued on (so callers can time it with events) */
int  odw_engine_stream(const odw_engine*, void** stream_out);

/* page-locked host memory for hit delivery (odw_trace_mc_host copies device->host asynchronously into it) */
int  odw_host_alloc(odw_engine*, uint64_t bytes, void** out);
void odw_host_free(odw_engine*, void* p);

int  odw_scene_create(odw_engine*, const odw_scene_desc*, odw_scene** out);
void odw_scene_destroy(odw_scene*);

int  odw_source_create(odw_engine*, const odw_source_desc*, odw_source** out);
void odw_source_destroy(odw_source*);

/* Monte-Carlo ('true' mode, point_source.py:659-679): rays [first_ray, first_ray+n_rays) of Philox stream
 * (seed, source_id).  Deterministic and independent of how the range is split across calls / GPUs. */
int  odw_trace_mc(odw_scene*, odw_source*, const odw_trace_cfg*, uint64_t seed,
                  uint64_t first_ray, uint64_t n_rays, odw_result** out);

/* Monte-Carlo trace with the hit list delivered straight into HOST arrays (page-locked memory recommended): the
 * range is traced in chunks and the device->host copy of chunk c overlaps the trace of chunk c+1.  Rows are appended
 * in chunk order (unsorted inside a chunk).  *n_hits_out = rows written; counts_out may be NULL.
 * cfg->hit_capacity = rows expected for the whole range (0 = two per ray): the per-chunk device hit lists are sized by
 * the same rows-per-ray ratio; hits that fit neither them nor host->capacity are counted in hits_dropped (ODW_EOVERFLOW). */
int  odw_trace_mc_host(odw_scene*, odw_source*, const odw_trace_cfg*, uint64_t seed, uint64_t first_ray, uint64_t n_rays,
                       const odw_hits_view* host, uint64_t* n_hits_out, odw_counts* counts_out);

/* Same draws as odw_trace_mc, returned instead of traced (parity of the sampler):
 * first_var[n] (theta | r), phi[n], origins[n][3], directions[n][3]; any may be NULL. */
int  odw_sample_mc(odw_source*, uint64_t seed, uint64_t first_ray, uint64_t n_rays,
                   double* first_var, double* phi, double* origins, double* directions);

/* Explicit ray list (fans, replay, parity): origins/directions [n][3] (directions need not be unit),
 * powers[n] (NULL = 1), ignore list applies to all rays. */
int  odw_trace_rays(odw_scene*, const odw_trace_cfg*, const double* origins, const double* directions,
                    const double* powers, const int32_t* ignored_groups, int32_t n_ignored,
                    uint64_t n_rays, odw_result** out);

int  odw_result_counts(const odw_result*, odw_counts* out);
/* copies min(hits, view->capacity) rows; *n_out = rows written.  sorted != 0 orders rows by (ray_index, bounce). */
int  odw_result_hits(const odw_result*, odw_hits_view* view, int sorted, uint64_t* n_out);
int  odw_result_histogram(const odw_result*, int32_t binning, double* bins_out /* [nu*nv] row-major (u, v) */);
/* device pointer of the fp64 bins of `binning` (for NCCL all-reduce by the caller); valid until destroy */
int  odw_result_histogram_device(const odw_result*, int32_t binning, void** dptr, uint64_t* n_bins);
/* per-ray end state for explicit lists / parity: n_segments[n] int32, final point [n][3], final power [n]; any may be NULL */
int  odw_result_ray_summary(const odw_result*, int32_t* n_segments, double* final_points, double* final_powers);
/* optical group each ray of an explicit list was travelling in when it ended, -1 = none (the medium of the last segment when
 * the ray escaped; with odw_hits_view.medium this gives the `media` list of *-rays.pkl, results_store.py:241-257) */
int  odw_result_ray_media(const odw_result*, int32_t* final_medium);
/* kernel time of the trace (CUDA events on the engine stream), milliseconds */
int  odw_result_kernel_ms(const odw_result*, double* ms);
void odw_result_destroy(odw_result*);

#ifdef __cplusplus
}
#endif
#endif /* ODW_H */
