"""Bind the calling process to the host cores (and, by first touch, the host memory) next to one GPU.

The host-facing step (ZoneVecEnv.step_host) moves ~40 bytes per env-step between a GPU and pinned host buffers, tens
of GB/s per GPU.  A process that `torchrun` started on an arbitrary core pins those buffers on whatever NUMA node it
happens to run on; with several ranks most of them then push their traffic across the socket interconnect.  Binding
each rank to the NUMA node its GPU hangs off, BEFORE the pinned buffers are allocated, keeps every rank's PCIe traffic
on its own node.  Host-side plumbing only (sysfs + sched_setaffinity); returns a description, never raises."""
import os
import subprocess


def bind_to_gpu_numa_node(device_index):
    try:
        import torch
        prop = torch.cuda.get_device_properties(device_index)
        bus = None
        if all(hasattr(prop, k) for k in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
            bus = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        else:
            q = subprocess.run(['nvidia-smi', '-i', str(device_index), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                               capture_output=True, text=True, timeout=20).stdout.strip().lower()
            if q:
                dom, rest = q.split(':', 1)
                bus = dom[-4:] + ':' + rest
        if not bus:
            return {'bound': False, 'why': 'no PCI bus id'}
        node_path = f'/sys/bus/pci/devices/{bus}/numa_node'
        if not os.path.exists(node_path):
            return {'bound': False, 'why': f'{node_path} missing', 'pci': bus}
        node = int(open(node_path).read().strip())
        if node < 0:
            return {'bound': False, 'why': 'no NUMA affinity reported', 'pci': bus}
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {'bound': False, 'why': 'node cores not in the allowed set', 'pci': bus, 'numa_node': node}
        os.sched_setaffinity(0, allowed)
        return {'bound': True, 'pci': bus, 'numa_node': node, 'cores': len(allowed)}
    except Exception as ex:                                    # never take a run down for a placement hint
        return {'bound': False, 'why': repr(ex)}
