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
