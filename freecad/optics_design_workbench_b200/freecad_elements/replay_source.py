'''
Replay light source: every hit stored by an earlier simulation becomes a new ray.

Mirrors ReplaySourceProxy of the reference (reference freecad_elements/replay_source.py:73-166): all `*-hits.pkl` files
below ReplayFromDir are read; hit i gives a ray with origin points[i], direction directions[i], power powers[i] and
wavelength data['wavelength'][i] if that column exists, else 1 (:107-111); origin and origin+direction go through the
source's global placement (:146-155).  The reference walks the files in random order and marks them consumed with flag
files so that its CPU workers share the stock; here the stock is one array, GPU ranks take disjoint index ranges of it
(the order of independent rays does not change any result).  When the stock is exhausted the simulation ends (:160-161).
'''

import glob
import os
import pickle

import numpy as np

from .point_source import RayBatch


class ReplayStock:
  def __init__(self, origins, directions, powers, wavelengths):
    self.origins, self.directions, self.powers, self.wavelengths = origins, directions, powers, wavelengths

  def __len__(self):
    return len(self.powers)

  def take(self, first, n):
    'rays [first, first+n) of the stock as RayBatches of one wavelength each (the engine traces a list at one wavelength)'
    lo, hi = min(first, len(self)), min(first+n, len(self))
    out = []
    wl = self.wavelengths[lo:hi]
    for w in np.unique(wl):
      sel = np.nonzero(wl == w)[0]+lo
      out.append(RayBatch(self.origins[sel], self.directions[sel], self.powers[sel], w, {}))
    return out


def load_stock(obj):
  'obj: source record with ReplayFromDir and gpM'
  folder = obj.get('ReplayFromDir')
  if not folder:
    raise RuntimeError(f"please set a replay directory for light source {obj.get('name')}")
  if not os.path.exists(folder):
    raise RuntimeError(f"selected replay directory of light source {obj.get('name')} does not seem to exist: {folder}")
  files = sorted(glob.glob(os.path.join(folder, '**', '*-hits.pkl'), recursive=True))
  if not files:
    raise RuntimeError(f"selected replay directory of light source {obj.get('name')} does not seem to contain any "
                       f"ray hit datafile: {folder}")
  P, D, W, L = [], [], [], []
  for f in files:
    with open(f, 'rb') as fh:
      data = pickle.load(fh)
    n = len(data['powers'])
    P.append(np.asarray(data['points'], dtype=np.float64).reshape(n, 3))
    D.append(np.asarray(data['directions'], dtype=np.float64).reshape(n, 3))
    W.append(np.asarray(data['powers'], dtype=np.float64).reshape(n))
    wl = np.ones(n)
    have = np.asarray(data.get('wavelength', []), dtype=np.float64)
    wl[:min(n, len(have))] = have[:n]
    L.append(wl)
  P, D, W, L = np.concatenate(P), np.concatenate(D), np.concatenate(W), np.concatenate(L)
  M = np.asarray(obj.get('gpM', np.eye(4)), dtype=np.float64).reshape(4, 4)
  p1 = P @ M[:3, :3].T + M[:3, 3]
  p2 = (P+D) @ M[:3, :3].T + M[:3, 3]
  return ReplayStock(np.ascontiguousarray(p1), np.ascontiguousarray(p2-p1), W, L)
