"""Multi-GPU sharding by stream: one process per GPU, independent streams dealt to ranks by size,
no data-path collective.  The reference's `DeflateFilesContainer.optimise` loop (cont/DeflateFilesContainer.java:
22-36) treats streams independently, so a partition of the list is a partition of the work.

torch.distributed is used for host-side plumbing only (rank/world discovery, gathering the per-rank result lists,
max-over-ranks timing); it is optional: without an initialised process group everything runs as world size 1.
"""
import heapq


def shard_streams(sizes, world_size):
    """Longest-processing-time greedy bin packing of stream indices onto ranks.  Returns a list (per rank) of index
    lists; every index appears exactly once; ties keep the original order, so the result is deterministic."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + sizes[i], r))
    for s in shards:
        s.sort()
    return shards


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


def rank_world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def optimise_sharded(buffers, merge_blocks=True, worker=None, gather=True):
    """Optimise a list of raw deflate streams across the ranks of the current process group.

    Every rank passes the same `buffers` (or at least the same list of sizes); rank r runs `worker` (default: the CUDA
    batch entry `deft4j_b200.optimise_batch`) on its shard.  With gather=True every rank receives the full result
    list in the original order (host-side all_gather_object); otherwise a rank gets its own results and None elsewhere.
    """
    if worker is None:
        from .deft import optimise_batch as worker
    rank, world = rank_world()
    shards = shard_streams([len(b) for b in buffers], world)
    mine = shards[rank]
    local = worker([buffers[i] for i in mine], merge_blocks) if mine else []
    out = [None] * len(buffers)
    if world == 1 or not gather:
        for i, r in zip(mine, local):
            out[i] = r
        return out
    gathered = [None] * world
    _dist().all_gather_object(gathered, list(zip(mine, local)))
    for part in gathered:
        for i, r in part:
            out[i] = r
    return out
