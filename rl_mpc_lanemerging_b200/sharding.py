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


def gather(values, device):
    """Every rank's list of floats, as a list of lists indexed by rank (diagnostics: per-rank step times and clocks)."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [[float(x) for x in o] for o in out]
    return [[float(x) for x in t]]


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
    (one process per GPU: the launch thread and the pinned-memory copies should not cross the socket interconnect).  Where the
    PCI topology is not exposed (numa_node = -1 in many containers and VMs) the allowed cores are split evenly between the local
    ranks instead, so that eight launch threads at least do not migrate or share cores.
    Returns (core list, "numa" | "even split"), or (None, reason) -- never raises: affinity is an optimisation, not a requirement."""
    import os
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except Exception as e:          # noqa: BLE001
        return None, f"no sched_getaffinity: {e!r}"

    def node_of(rank):
        props = torch.cuda.get_device_properties(rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"{sysfs}/bus/pci/devices/{bdf}/numa_node") as f:
            return int(f.read())

    try:
        node = node_of(local_rank)
        if node < 0:
            raise LookupError("numa_node = -1")
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as f:
            cpus = [c for c in _parse_cpulist(f.read()) if c in allowed]
        peers = [r for r in range(local_world) if node_of(r) == node]       # ranks whose GPUs sit on the same node share its cores
        share = max(len(cpus) // max(len(peers), 1), 1)
        k = peers.index(local_rank)
        mine, how = (cpus[k * share:(k + 1) * share] or cpus), "numa"
    except Exception as e:          # noqa: BLE001
        if not torch.cuda.is_available() or local_world < 1 or len(allowed) < 2 * local_world:
            return None, f"topology unreadable ({e!r}) and too few cores to split"
        share = len(allowed) // local_world
        mine, how = allowed[local_rank * share:(local_rank + 1) * share], "even split"
    if not mine:
        return None, "empty core list"
    try:
        os.sched_setaffinity(0, mine)
    except Exception as e:          # noqa: BLE001
        return None, f"sched_setaffinity failed: {e!r}"
    return mine, how
