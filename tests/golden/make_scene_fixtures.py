'''
Exports the reference's benchmark scenes (and a few test scenes) through the headless FCStd importer
into tests/golden/scenes/<name>.npz.  /root/reference does not exist on the GPU box, so bench.py,
smoke() and the `-m gpu` tests load these fixtures; run this script in the build container after any
change to scene_export/.  The fixtures hold OUR flat scene description (faces, trims, groups, source
properties, settings) — not the FCStd archives.
'''
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from freecad.optics_design_workbench_b200.simulation.setup import PreparedSimulation

SCENES = {
  'minimal': 'benchmark/minimal.FCStd',
  'lensesAndMirrors': 'benchmark/lensesAndMirrors.FCStd',
  'lensesAndMirrorsSequential': 'benchmark/lensesAndMirrorsSequential.FCStd',
  'hugeArray': 'benchmark/hugeArray.FCStd',
  # reference test/21-simulation-modes: surface source (Box001.Face5, cos(theta)**2) -> sphere lens -> absorber box, sequential mode
  'surfaceSourceTest21': 'test/21-simulation-modes/main.FCStd',
  # reference test/50-old-tests: transmission/reflection grating and a Gaussian beam on a box detector
  'grating': 'test/50-old-tests/grating.FCStd',
  'gaussian': 'test/50-old-tests/gaussian.FCStd',
  'mirrorDiffuse': 'test/50-old-tests/mirror-diffuse.FCStd',     # Lambert-like diffuse mirror (stochastic surface model)
  'gettingStarted': 'examples/1-getting-started/GettingStarted.FCStd',
}

def main():
  for name, rel in SCENES.items():
    sim = PreparedSimulation.from_fcstd(os.path.join('/root/reference', rel))
    out = os.path.join(HERE, 'scenes', name+'.npz')
    sim.save_fixture(out)
    print(name, sim.scene.summary(), 'skipped faces:', len(sim.info['skipped']), '->', os.path.getsize(out), 'bytes')

if __name__ == '__main__':
  main()
