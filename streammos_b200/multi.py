"""Multi-GPU plumbing. The hot path shards by INDEPENDENT scan streams (sequences): a stream is strictly
sequential (short-term memory recurrence, 8-scan voting window), streams share nothing. One process per
GPU, each owning whole sequences and their memories in its own HBM; no collective on the data path —
torch.distributed is used only for the start barrier and the max-over-ranks of the timing."""


def sequences_of_rank(sequences, rank, world, lengths=None):
    """Static assignment of whole sequences to ranks. With `lengths` (scans per sequence) the assignment is
    longest-first onto the least-loaded rank (SemanticKITTI sequences differ 10x in length); otherwise
    round-robin."""
    if lengths is None:
        return [s for i, s in enumerate(sequences) if i % world == rank]
    load = [0] * world
    owner = {}
    for s in sorted(sequences, key=lambda s: (-lengths[s], str(s))):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[s] = r
        load[r] += lengths[s]
    return [s for s in sequences if owner[s] == rank]


def stream_seed(rank, stream_index):
    """Seed of the synthetic stream `stream_index` of `rank` (distinct streams per rank)."""
    return 1000 * rank + stream_index


def aggregate_scans_per_second(per_rank_steps, max_ms_total, world):
    """Whole-job throughput: every rank processed `per_rank_steps` scans in (at most) `max_ms_total` ms."""
    return world * per_rank_steps / (max_ms_total / 1e3)


def gpu_local_cpus(pci_bus_id, sysfs="/sys/bus/pci/devices"):
    """CPUs on the NUMA node a GPU's PCIe root hangs off (`local_cpulist` of its sysfs entry), as a sorted list, or
    [] when the platform does not say. `pci_bus_id`: 'domain:bus:device.function' as CUDA / nvidia-smi print it."""
    import os
    dom, bus, rest = pci_bus_id.strip().lower().split(":")
    path = os.path.join(sysfs, "%04x:%s:%s" % (int(dom, 16), bus, rest), "local_cpulist")
    try:
        text = open(path).read().strip()
    except OSError:
        return []
    return parse_cpulist(text)


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the kernel's cpulist format)."""
    cpus = set()
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return sorted(cpus)


def share_of_cpus(cpus, slot, slots):
    """Contiguous share `slot` of `slots` of a CPU list (ranks whose GPUs sit on the same NUMA node split its cores
    instead of all crowding the first ones). Never empty if `cpus` is not."""
    if not cpus or slots <= 1:
        return list(cpus)
    per = max(1, len(cpus) // slots)
    lo = (slot % slots) * per
    out = cpus[lo:lo + per] if slot % slots < slots - 1 else cpus[lo:]
    return out or list(cpus)


def pin_rank_to_gpu(local_rank, world_local):
    """Bind this process (and the pinned host buffers it allocates afterwards: first touch) to the cores next to its
    GPU. With one process per GPU feeding 8 H2D streams from one host, a rank running on the far socket pays the
    inter-socket link on every submit and every byte of its pinned buffers. Returns the CPU list used, or None when
    nothing was changed (single rank, no sysfs entry, affinity not permitted)."""
    import os
    if world_local <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        cpus = gpu_local_cpus("%04x:%02x:%02x.0" % (dom, bus, dev))
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        # ranks on the same node: split the node's cores by local rank parity within the node (best effort — the
        # exact sibling count is unknown without querying every GPU, so use world_local / 2 per node)
        mine = share_of_cpus(cpus, local_rank, max(1, world_local // 2))
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None
