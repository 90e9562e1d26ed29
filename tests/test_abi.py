'''
The C-ABI library loads and exports every symbol include/odw.h declares; without a GPU the product path
fails loudly instead of falling back to anything on the CPU.  No compute calls here.
'''
import ctypes as C
import os
import re

import numpy as np
import pytest

from freecad.optics_design_workbench_b200 import engine, _abi
from freecad.optics_design_workbench_b200.scene_export import scene as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
  text = open(os.path.join(ROOT, 'include', 'odw.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(odw_[a-z_0-9]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
  lib = engine.load_library()
  names = declared_functions()
  assert len(names) >= 19
  for n in names:
    assert hasattr(lib, n), f'{n} declared in include/odw.h but not exported by libodw_b200.so'
  assert set(names) == set(engine.EXPORTS)
  assert lib.odw_abi_version() == 2


def test_struct_layouts_match_header():
  'numpy dtypes / ctypes structs mirror the C structs: sizes as the C compiler sees include/odw.h (via the oracle build)'
  import ctypes
  from oracle import Oracle
  lib = Oracle().lib
  lib.oracle_sizeof.restype = ctypes.c_int
  csize = lambda which: lib.oracle_sizeof(which)
  assert sc.FACE_DTYPE.itemsize == csize(0) == 224 and sc.SEG_DTYPE.itemsize == csize(1) == 48
  assert sc.SHELL_DTYPE.itemsize == csize(2) == 64 and sc.GROUP_DTYPE.itemsize == csize(3) == 80
  assert C.sizeof(_abi.SceneDesc) == csize(4)
  assert C.sizeof(_abi.SourceDesc) == csize(5)
  assert C.sizeof(_abi.Binning) == csize(6)
  assert C.sizeof(_abi.TraceCfg) == csize(7)
  assert C.sizeof(_abi.Counts) == csize(8) == 64
  assert C.sizeof(_abi.HitsView) == csize(9) == 80


@pytest.mark.skipif(os.environ.get('ODW_EXPECT_GPU') == '1', reason='GPU box')
def test_no_cpu_fallback_without_device():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a CUDA device is present')
  with pytest.raises(engine.EngineError) as e:
    engine.Engine(0)
  assert e.value.code == _abi.ODW_ENODEVICE
  assert 'no CPU fallback' in str(e.value)


def test_product_package_does_not_import_oracle():
  'the oracle is test infrastructure: nothing under the product package may import it'
  pkg = os.path.join(ROOT, 'freecad', 'optics_design_workbench_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
        text = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f'{f} imports oracle'
        assert 'odw_oracle' not in text, f'{f} references the oracle library'


def test_kernel_instance_selection():
  '''
  Which instance of the trace kernel a launch gets (csrc/odw_kernels.cu pick_feat; FEAT_* bits: 1 gratings / scatter / absorption /
  aspheres, 2 surface source, 4 sequential mode, 8 device binning): the lean one only when nothing is needed, the sequential
  one for sequential mode alone, the one without FEAT_EXT for surface sources / device binning in scenes of ideal surfaces,
  the full one otherwise and for BVH scenes; explicit ray lists get the one without FEAT_EXT when they can.  Host logic, no GPU call.
  '''
  import ctypes as C
  from freecad.optics_design_workbench_b200 import engine
  L = engine.load_library()
  L.odw_trace_instance.restype = C.c_int
  L.odw_trace_instance.argtypes = [C.c_bool, C.c_bool, C.c_int]
  assert L.odw_trace_instance(True, False, 0) == 0
  assert L.odw_trace_instance(True, False, 4) == 4
  for need in (2, 8, 6, 12, 14):
    assert L.odw_trace_instance(True, False, need) == 14, need    # surface source / binning without gratings, scatter, ...: everything but FEAT_EXT
  for need in (1, 5, 3, 9, 15):
    assert L.odw_trace_instance(True, False, need) == 15, need
  for need in (0, 4, 15):
    assert L.odw_trace_instance(False, False, need) == (15 if need & 1 else 14)    # explicit ray lists (fans, replays, pre-sampled surface sources)
    assert L.odw_trace_instance(True, True, need) == 15         # BVH scenes
    assert L.odw_trace_instance(False, True, need) == 15
