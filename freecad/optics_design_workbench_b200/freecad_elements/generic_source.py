'''
The call site the engine replaces: GenericSourceProxy.runSimulationIteration
(reference freecad_elements/generic_source.py:51-146).

The reference generates N ray objects, walks each through Ray.traceRay and hands every hit to
obj.Proxy.onRayHit -> store.addRayHit.  Here one engine call traces the whole batch on the GPU and the
hits come back as arrays, grouped per optical group and appended to the store in one go
(optical_group.py:206-209 -> results_store.py:641-648 semantics: only groups with RecordHits).

Same name, same keyword arguments, same side effects on `store` (hits, totalTracedRays).  `draw=True`
(GUI ray drawing, :102-138): the rays are traced on the GPU with every intersection recorded and their polylines go to
a drawing back end (freecad_elements/ray_drawing.py: the reference's Part.makeLine / RaySegment calls inside FreeCAD, a
recording back end elsewhere).
'''

import numpy as np

from . import point_source
from ..simulation import results_store


class GenericSourceProxy:
  '''
  One instance per light source, bound to a SimulationContext (simulation/simulation_loop.py), which plays
  the role of the reference's module-level singletons (simulatingDocument(), find.activeSimulationSettings()).
  '''
  def __init__(self, context, source_index):
    self.context = context
    self.index = source_index

  @property
  def record(self):
    return self.context.sim.source_records[self.index]

  # -- ray generation (PointSourceProxy._generateRays) --------------------------------------------
  def _generateRays(self, obj, mode, **kwargs):
    if mode == 'fans' and obj.get('proxy') == 'SurfaceSourceProxy':
      from . import surface_source
      emit = obj.get('emit')
      if emit is None or not len(emit.faces):
        raise NotImplementedError(f"surface source {obj['name']}: no emitting faces ({obj.get('emit_error') or 'ActiveSurfaces empty'})")
      return surface_source.generate_fan_rays(obj, emit, self.context.sim.settings.get('DistanceTolerance', 1e-6))
    if mode == 'fans':
      return point_source.generate_fan_rays(obj, obj['gpM'], max_fan_count=kwargs.get('maxFanCount', np.inf),
                                            max_rays_per_fan=kwargs.get('maxRaysPerFan', np.inf))
    raise ValueError(f'unexpected ray placement mode {mode} for host-side ray generation')

  def _pseudo_rays(self, obj, iterations):
    '''
    mode 'pseudo' (point_source.py:671-679): per iteration RaysPerIteration*scale pseudo-random (theta|r, phi) pairs from
    drawPseudo, generated on the host (a sequential, data-dependent thinning procedure) and traced as an explicit list.
    This rank's share of the iterations is drawn from a numpy stream derived from (seed, source id, first iteration).
    '''
    from ..distributions import sampler_tables as st
    ctx = self.context
    n_iter = point_source.rays_per_iteration(obj, ctx.sim.settings)
    first_it, my_its = ctx.claim_rays(('pseudo', self.index), iterations)
    sa = ctx.sim.source_args(self.index)
    expr, _ = st.point_source_density(obj['PowerDensity'], float(obj['FocalLength']))
    rng = np.random.default_rng([ctx.seed & 0xFFFFFFFF, ctx.seed >> 32, int(obj.get('source_id', self.index)), first_it])
    firsts, phis = [], []
    for _ in range(my_its):
      f, p = st.draw_pseudo(sa.tables, expr, n_iter, rng)
      firsts.append(f); phis.append(p)
    firsts = np.concatenate(firsts) if firsts else np.zeros(0)
    phis = np.concatenate(phis) if phis else np.zeros(0)
    return point_source.make_rays(obj, obj['gpM'], firsts, phis)

  # -- the iteration --------------------------------------------------------------------------------
  def runSimulationIteration(self, obj=None, *, mode, draw=False, store=False, returnInitialConditions=False,
                             useInitialConditions=None, iterations=1, drawBackend=None, sourceObject=None, **kwargs):
    '''
    mode 'true'  : `iterations` Monte-Carlo iterations of RaysPerIteration*RaysPerIterationScale rays each in ONE
                   engine call (odw_trace_mc), drawn on the device from the Philox stream (seed, source id) at
                   the context's next free global ray indices; in a multi-GPU run this rank traces its shard.
    mode 'fans'  : the deterministic fan list (or `useInitialConditions`, a RayBatch) through odw_trace_rays.
    Returns the engine counters of this call (the reference returns None), or the RayBatch if
    returnInitialConditions.
    '''
    ctx = self.context
    obj = obj if obj is not None else self.record
    kind = obj.get('proxy', 'PointSourceProxy')
    if draw:
      # displayed rays (generic_source.py:102-138): traced as an explicit list with every intersection recorded, then drawn
      from . import ray_drawing
      backend = drawBackend if drawBackend is not None else ray_drawing.default_backend()
      if useInitialConditions is not None:
        batch = useInitialConditions
      elif mode in ('fans', 'multicorefans'):
        batch = self._generateRays(obj, mode='fans', **kwargs)
      elif mode in ('pseudo', 'singlepseudo'):
        batch = self._pseudo_rays(obj, int(iterations))
      elif kind == 'ReplaySourceProxy':
        raise NotImplementedError('drawing the rays of a replay source')
      else:
        first, n = ctx.claim_rays(self.index, point_source.rays_per_iteration(obj, ctx.sim.settings)*int(iterations))
        smp = ctx.device_source(self.index).sample(ctx.seed, first, n)
        batch = point_source.RayBatch(smp['origins'], smp['directions'], np.ones(n), float(obj['Wavelength']), {})
      counts, rays = self._trace_explicit(obj, batch, store, record_rays=bool(store and obj.get('RecordRays', False)), return_rays=True)
      ray_drawing.draw_rays(sourceObject if sourceObject is not None else obj, rays, obj.get('gpM', np.eye(4)), backend)
      return counts
    if kind not in ('PointSourceProxy', 'SurfaceSourceProxy', 'ReplaySourceProxy'):
      raise NotImplementedError(f"light source kind {kind} is not handled by the engine yet")
    if kind == 'ReplaySourceProxy':
      return self._replay_iteration(obj, mode, int(iterations), store, returnInitialConditions)

    if useInitialConditions is not None or mode in ('fans', 'multicorefans'):
      batch = useInitialConditions if useInitialConditions is not None else self._generateRays(obj, mode='fans', **kwargs)
      if returnInitialConditions:
        return batch
      return self._trace_explicit(obj, batch, store, record_rays=bool(store and obj.get('RecordRays', False)))
    if mode in ('pseudo', 'singlepseudo'):
      if kind != 'PointSourceProxy':
        raise NotImplementedError('pseudo-random mode is implemented for point sources only')
      batch = self._pseudo_rays(obj, int(iterations))
      if returnInitialConditions:
        return batch
      return self._trace_explicit(obj, batch, store, record_rays=bool(store and obj.get('RecordRays', False)))
    if mode not in ('true', 'singletrue'):
      raise ValueError(f'unexpected ray placement mode {mode}')
    if returnInitialConditions:
      raise NotImplementedError('Monte-Carlo rays are drawn on the device; use DeviceSource.sample for their initial conditions')
    if store and obj.get('RecordRays', False) and kind == 'PointSourceProxy':
      return self._trace_monte_carlo_recording_rays(obj, int(iterations), store)
    return self._trace_monte_carlo(obj, int(iterations), store)

  # -- replay source (reference freecad_elements/replay_source.py:73-166) --------------------------------
  def _replay_iteration(self, obj, mode, iterations, store, returnInitialConditions):
    from . import replay_source
    if mode in ('fans', 'multicorefans'):
      return None                                        # :132-135 a replay source places no fans
    stock = self.context.replay_stock(self.index, lambda: replay_source.load_stock(obj))
    n_total = point_source.rays_per_iteration(obj, self.context.sim.settings)*iterations
    first, n = self.context.claim_rays(self.index, n_total)
    batches = stock.take(first, n)
    if returnInitialConditions:
      return batches
    counts = None
    for batch in batches:
      c = self._trace_explicit(obj, batch, store)
      counts = c if counts is None else {k: counts[k]+c[k] for k in counts}
    if first+n >= len(stock) and first < len(stock) or not batches:
      from ..simulation.simulation_loop import SimulationEnded
      raise SimulationEnded(f'replay light source {obj["name"]} ran out of rays')     # :160-161
    return counts

  # -- engine calls -----------------------------------------------------------------------------------
  def _store_hits(self, obj, hits, store, metadata_of, borrowed=False):
    '''
    append the hit arrays to the store, one entry per optical group (file per (source, object)).  borrowed: the arrays
    are views of the engine's page-locked delivery buffers (valid until the next engine call); hits of a single group
    are handed over as they are, without selecting or copying.
    '''
    scene = self.context.sim.scene
    group = hits['group']
    keys = self.context.sim.settings.get('store_hit_keys', [])
    if not len(group):
      return
    if getattr(group, 'strides', (1,))[0] == 0:                   # one recording group, broadcast instead of a column: nothing to count
      present = np.array([int(group[0])])
    else:
      present = np.flatnonzero(np.bincount(group, minlength=len(scene.group_names)))
    for gi in present:
      if len(present) == 1:
        sel = slice(None)
      else:
        sel = np.flatnonzero(group == gi)
      md = metadata_of(hits['ray_index'][sel], keys) if keys else {}
      store.addRayHits(results_store.named((obj['name'], obj['label'])),
                       results_store.named((scene.group_names[gi], scene.group_labels[gi])),
                       hits['points'][sel], hits['directions'][sel], hits['powers'][sel], hits['is_entering'][sel], md,
                       borrowed=borrowed and len(present) == 1)

  def _ray_dicts(self, obj, batch, hits, summary):
    '''
    Per-ray polylines in the reference's *-rays.pkl form (SimulationResultsSingleRay.dump, results_store.py:241-257):
    points = segment start points + the end of the last segment, powers = power at the start of each segment, media =
    Name of the optical group each segment runs through (None = vacuum).  Needs every intersection recorded.
    '''
    scene = self.context.sim.scene
    names = scene.group_names
    refl = np.where(scene.groups['optical_type'] == 0, scene.groups['reflectivity'], 1.0)      # Mirror: power *= Reflectivity
    refl = np.where(scene.groups['optical_type'] == 3, 0.0, refl)                              # Absorber: power = 0
    order = np.lexsort((hits['bounce'], hits['ray_index']))
    ray = hits['ray_index'][order].astype(np.int64)
    starts = np.searchsorted(ray, np.arange(len(batch)+1))
    out = []
    for r in range(len(batch)):
      sel = order[starts[r]:starts[r+1]]
      n_seg = int(summary['n_segments'][r])
      pts = [batch.origins[r]] + list(hits['points'][sel])
      media = [None if m < 0 else names[m] for m in hits['medium'][sel]]
      powers = [batch.powers[r]] + list(hits['powers'][sel]*refl[hits['group'][sel]])
      if n_seg > len(sel):                                            # the ray escaped: its last segment has no hit
        pts.append(summary['final_points'][r])
        fm = int(summary['final_media'][r])
        media.append(None if fm < 0 else names[fm])
      else:
        powers = powers[:-1]
      if n_seg:
        out.append(dict(points=np.array(pts), powers=np.array(powers[:n_seg]), media=media))
    return out

  def _trace_explicit(self, obj, batch, store, record_rays=False, return_rays=False):
    'return_rays: also return the per-ray polylines (for drawing); every intersection is recorded then, stored or not'
    ctx = self.context
    want_rays = bool(return_rays or (store and record_rays))
    cfg = ctx.cfg(obj, store_hits=bool(store or return_rays), record_all_hits=want_rays,
                  hit_capacity=max(1024, len(batch)*int(ctx.sim.settings['MaxIntersections'])),
                  wavelength=batch.wavelength, scatter_seed=ctx.seed,
                  max_ray_length=float(ctx.sim.settings['MaxRayLength'])*float(obj.get('MaxRayLengthScale', 1.0)),
                  max_intersections=int(float(ctx.sim.settings['MaxIntersections'])*float(obj.get('MaxIntersectionsScale', 1.0))))
    with ctx.device_scene.trace_rays(cfg, batch.origins, batch.directions, batch.powers, ignored=obj.get('ignored', ())) as res:
      counts = res.counts
      hits = res.hits(sort=True) if (store or return_rays) else None
      summary = res.ray_summary() if want_rays else None
    rays = self._ray_dicts(obj, batch, hits, summary) if want_rays else None
    if store and record_rays:
      store.addRays(results_store.named((obj['name'], obj['label'])), rays)
    if want_rays and hits is not None:
      keep = ctx.sim.scene.groups['record_hits'][hits['group']] != 0        # onRayHit only stores RecordHits groups
      hits = {k: v[keep] for k, v in hits.items()}
    if store:
      def metadata_of(ray_index, keys):
        idx = ray_index.astype(np.int64)
        md = {}
        for k in keys:
          name = k[0].lower()+k[1:]                       # StoreHitInitPoint -> initPoint
          if name == 'initPoint': md[name] = batch.origins[idx]
          elif name == 'initDirection': md[name] = batch.directions[idx]
          elif name == 'initPower': md[name] = batch.powers[idx]
          elif name == 'initWavelength': md[name] = np.full(len(idx), batch.wavelength)
          elif name in batch.metadata: md[name] = batch.metadata[name][idx]
        return md
      self._store_hits(obj, hits, store, metadata_of)
      store.incrementRayCount(len(batch))
    return (counts, rays) if return_rays else counts

  def _trace_monte_carlo_recording_rays(self, obj, iterations, store):
    '''
    RecordRays (generic_source.py:80-82,96-100): the same Monte-Carlo rays (same Philox stream, odw_sample_mc) as an explicit
    list, traced with every intersection recorded so that the per-ray polylines can be written next to the hits.
    '''
    ctx = self.context
    first, n = ctx.claim_rays(self.index, point_source.rays_per_iteration(obj, ctx.sim.settings)*iterations)
    s = ctx.device_source(self.index).sample(ctx.seed, first, n)
    finite = np.isfinite(float(obj.get('FocalLength', 0)))
    md = dict(initPhi=s['phi'], initTheta=s['first'] if finite else np.full(n, np.nan))
    batch = point_source.RayBatch(s['origins'], s['directions'], np.ones(n), float(obj['Wavelength']), md)
    return self._trace_explicit(obj, batch, store, record_rays=True)

  def _trace_monte_carlo(self, obj, iterations, store):
    ctx = self.context
    n_iter_rays = point_source.rays_per_iteration(obj, ctx.sim.settings)
    n_total = n_iter_rays*iterations
    first, n = ctx.claim_rays(self.index, n_total)        # this rank's shard of the next n_total global ray indices
    dsrc = ctx.device_source(self.index)
    if not store:
      with ctx.device_scene.trace_mc(dsrc, ctx.cfg(obj, store_hits=False), ctx.seed, first, n) as res:
        return res.counts
    # Hit delivery straight into page-locked host arrays (odw_trace_mc_host: the device->host copy of one chunk overlaps
    # the trace of the next) with only the columns the result files need: the group column only when more than one
    # optical group records hits, the ray index only for StoreHit* metadata.
    scene = ctx.sim.scene
    recording = np.nonzero(scene.groups['record_hits'])[0]
    keys = ctx.sim.settings.get('store_hit_keys', [])
    columns = ['points', 'directions', 'powers', 'is_entering'] + (['group'] if len(recording) != 1 else []) + (['ray_index'] if keys else [])
    capacity = max(1024, 2*n)
    # a ray records at most one hit per intersection: that bounds the retries below
    ceiling = max(1024, n*(int(float(ctx.sim.settings['MaxIntersections'])*float(obj.get('MaxIntersectionsScale', 1.0)))+1))
    while True:
      arrays, view = ctx.pinned_hits(capacity, tuple(columns))
      # hit_capacity = rows expected for the whole range: the engine sizes its per-chunk device lists by the same ratio
      counts, got = ctx.device_scene.trace_mc_host(dsrc, ctx.cfg(obj, store_hits=True, hit_capacity=capacity), ctx.seed, first, n, view)
      if not counts['hits_dropped']:
        break
      if capacity >= ceiling:
        raise RuntimeError(f"{counts['hits_dropped']} hits dropped although the hit buffers hold one row per possible intersection")
      # more recorded hits than rows (transparent detectors record two hits per pass): the same ray range again with
      # room for all of them — the Philox stream makes the repeat identical
      capacity = min(ceiling, max(4*capacity, int(counts['hits'])+1024))
    hits = {k: v[:got] for k, v in arrays.items()}
    if 'group' not in hits:
      hits['group'] = np.broadcast_to(np.int32(recording[0]), (got,))         # one recording group: no column, no memory
    if 'ray_index' not in hits:
      hits['ray_index'] = np.broadcast_to(np.uint64(0), (got,))
    if store:
      def metadata_of(ray_index, keys):
        s = dsrc.sample(ctx.seed, first, n)               # same Philox stream -> the rays' initial conditions
        idx = (ray_index-np.uint64(first)).astype(np.int64)
        finite = np.isfinite(float(obj.get('FocalLength', 0)))
        md = {}
        for k in keys:
          name = k[0].lower()+k[1:]
          if name == 'initPoint': md[name] = s['origins'][idx]
          elif name == 'initDirection': md[name] = s['directions'][idx]
          elif name == 'initPower': md[name] = np.ones(len(idx))
          elif name == 'initWavelength': md[name] = np.full(len(idx), float(obj['Wavelength']))
          elif name == 'initPhi': md[name] = s['phi'][idx]
          elif name == 'initTheta': md[name] = s['first'][idx] if finite else np.full(len(idx), np.nan)
        return md
      self._store_hits(obj, hits, store, metadata_of, borrowed=True)          # views of the page-locked delivery buffers
      store.incrementRayCount(n)
    return counts
