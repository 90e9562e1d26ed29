'''Developer script (run under torchrun on an N-GPU box): device->host copy bandwidth per rank, alone and all ranks at once,
with the process where the launcher put it and again after moving it (and its page-locked buffer) to the NUMA node of its GPU.
Tells whether the end-to-end number at N > 1 is limited by where the pinned host buffers live.'''
import os, sys, time
import torch
import torch.distributed as dist


def node_of_gpu(i):
  p = torch.cuda.get_device_properties(i)
  try:
    bdf = f'{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0'
    with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as f:
      return bdf, int(f.read())
  except Exception as e:
    return repr(e), -1


def cpus_of_node(n):
  try:
    with open(f'/sys/devices/system/node/node{n}/cpulist') as f:
      out = set()
      for part in f.read().strip().split(','):
        a, _, b = part.partition('-')
        out.update(range(int(a), int(b or a)+1))
      return out
  except Exception:
    return set()


def d2h(dev_buf, host_buf, reps=4):
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  host_buf.copy_(dev_buf, non_blocking=True)
  torch.cuda.synchronize()
  e0.record()
  for _ in range(reps):
    host_buf.copy_(dev_buf, non_blocking=True)
  e1.record()
  torch.cuda.synchronize()
  return reps*dev_buf.numel()*dev_buf.element_size()/(e0.elapsed_time(e1)*1e-3)/1e9


def main():
  rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('gloo')
  bar = (lambda: dist.barrier()) if world > 1 else (lambda: None)
  n = 1 << 28
  dev = torch.empty(n, dtype=torch.float32, device='cuda')
  bdf, node = node_of_gpu(local)
  aff = sorted(os.sched_getaffinity(0))
  for phase in ('as launched', 'bound to the GPU node'):
    if phase != 'as launched':
      cpus = cpus_of_node(node) & set(os.sched_getaffinity(0)) if node >= 0 else set()
      if cpus:
        os.sched_setaffinity(0, cpus)
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    host.zero_()
    alone = 0.0
    for r in range(world):
      bar()
      if r == rank:
        alone = d2h(dev, host)
    bar()
    together = d2h(dev, host, reps=8)
    bar()
    print(f'[{phase}] rank {rank} gpu {bdf} numa {node} cpus {len(os.sched_getaffinity(0))} (was {len(aff)}): '
          f'alone {alone:.1f} GB/s, all ranks at once {together:.1f} GB/s', flush=True)
    del host
    bar()
  if rank == 0:
    os.system('nvidia-smi topo -m 2>&1 | head -30; lscpu | grep -i "numa\\|socket\\|model name" ')


if __name__ == '__main__':
  main()
