"""Stream -> GPU partition: streams are independent, so the path shards with no
data-path collective.  Contiguous blocks: rank r of W owns streams [floor(r*N/W), floor((r+1)*N/W))."""


def shard_range(n_streams, rank, world):
    lo = (n_streams * rank) // world
    hi = (n_streams * (rank + 1)) // world
    return lo, hi


def shard_of(stream, n_streams, world):
    """Inverse of shard_range: the rank owning `stream`."""
    r = min(world - 1, (stream * world) // max(n_streams, 1))
    while stream < shard_range(n_streams, r, world)[0]:
        r -= 1
    while stream >= shard_range(n_streams, r, world)[1]:
        r += 1
    return r
