"""Episode sharding for multi-GPU runs: the planner path needs no exchange (independent episodes), so ranks are
replicas that own contiguous blocks of global episode ids; only timings / counters are reduced."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous block of episodes owned by `rank`: [lo, hi).  Blocks differ in size by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max(values, device):
    """Max over ranks of a list of floats (timings): the slowest rank defines the step time."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def reduce_sum(values, device):
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_local_cores(local_rank: int, local_world: int, sysfs: str = "/sys"):
    """Pin this process to the cores of the NUMA node its GPU hangs off, split evenly between the ranks that share the node
    (one process per GPU: the launch thread and the pinned-memory copies should not cross the socket interconnect).
    Returns the core list it bound to, or None when the topology cannot be read (containers without sysfs PCI entries) --
    never raises: affinity is an optimisation, not a requirement."""
    import os
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"{sysfs}/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as f:
            cpus = [c for c in _parse_cpulist(f.read()) if c in os.sched_getaffinity(0)]
        # ranks whose GPUs sit on the same node share its cores
        peers = []
        for r in range(local_world):
            q = torch.cuda.get_device_properties(r)
            qbdf = f"{q.pci_domain_id:04x}:{q.pci_bus_id:02x}:{q.pci_device_id:02x}.0"
            with open(f"{sysfs}/bus/pci/devices/{qbdf}/numa_node") as f:
                if int(f.read()) == node:
                    peers.append(r)
        share = max(len(cpus) // max(len(peers), 1), 1)
        k = peers.index(local_rank)
        mine = cpus[k * share:(k + 1) * share] or cpus
        if not mine:
            return None
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:               # noqa: BLE001
        return None
