"""Multi-GPU plumbing: shard subjects over ranks, gather per-shard feature tables.

The hot path shards naturally -- windows never span series and day segments never span subjects
 -- so every rank runs the single-GPU kernels on its own subjects with NO data-path
collective; the only exchange is one gather of the (small, fixed-width) feature tables, over NCCL on
GPUs (NVLink 5 / NVSwitch) or gloo in the CPU tests.  One process per GPU, torch.distributed.
"""


def device_numa_node(device_index):
    """NUMA node of the GPU's PCIe root, from sysfs via the bus id torch reports (no pynvml needed); None when the
    box has a single node or does not say (numa_node = -1, containers without /sys)."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(int(device_index))
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_host_to_device_numa(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers allocated
    afterwards (first touch) live next to the GPU's PCIe root: with one process per GPU, eight ranks pulling 54 GB/s
    each otherwise meet on the cross-socket link.  Best effort: returns the node id, or None when the topology cannot
    be read or the box has one node (then there is nothing to bind)."""
    import os
    try:
        node = device_numa_node(device_index)
        if node is None:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def shard_range(n_units, rank, world):
    """Contiguous balanced block of unit ids for ``rank``: [start, stop)."""
    base, rem = divmod(int(n_units), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_units, world):
    return [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]


def gather_tables(local, n_units, group=None, dst=None):
    """All-gather (dst=None) or gather-to-``dst`` of per-rank tables whose leading dimension indexes the
    rank's units (``shard_range`` order).  Ranks may hold different unit counts: rows are padded to the
    largest shard for the collective and trimmed afterwards.  Returns the full [n_units, ...] table
    (on every rank, or on ``dst`` only -- other ranks get None)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_units, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d units, expected %d" % (rank, local.shape[0], sizes[rank]))
    pad = max(sizes)
    buf = local
    if local.shape[0] != pad:
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
    buf = buf.contiguous()
    if dst is None:
        full = torch.empty((world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, buf, group=group)
        parts = [full[r * pad:r * pad + sizes[r]] for r in range(world)]
        return torch.cat(parts, dim=0)
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([outs[r][:sizes[r]] for r in range(world)], dim=0)


def sharded_window_table(make_series, n_subjects, wsize, wstep, features, group=None, **kw):
    """Convenience driver: ``make_series(first, count)`` yields this rank's [count * series_per_subject, len]
    device tensor; every rank extracts its table and the tables are all-gathered.  Used by the tests and
    as a template for callers (bench.py keeps its tables sharded: the full config-3 table is 23 GB)."""
    import torch.distributed as dist
    from . import engine
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = shard_range(n_subjects, rank, world)
    x = make_series(a, b - a)
    per_subject = x.shape[0] // max(1, b - a)
    tab = engine.window_table(x, wsize, wstep, features, **kw)
    tab = tab.reshape((b - a, per_subject) + tuple(tab.shape[1:]))
    return gather_tables(tab, n_subjects, group)
