"""shard.py -- multi-GPU decomposition of the (T) energy.

Units are occupied triples (lexicographic i<j<k in SYMMETRIC mode, all ordered (i,j,k) in
LITERAL mode); they are independent and equal-cost, so rank r of W evaluates the triples
{r, r+W, r+2W, ...}: a static round-robin that is identical for every GPU count.  The only
exchange step is one all-reduce.  Each rank fills its own entries of a zero vector with one
slot per triple; summing such vectors is exact (x + 0 + ... + 0), so after the all-reduce
every rank holds the same per-triple energies and adds them in index order: the result is
bit-identical for 1, 2, 4 or 8 ranks.
"""


def rank_range(rank, world):
    """(first, stride) of the triples owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank, world


def rank_count(ntriples, rank, world):
    first, stride = rank_range(rank, world)
    return 0 if first >= ntriples else (ntriples - first + stride - 1) // stride


def reduce_triple_energies(e_triples, dist=None):
    """All-reduce (sum) the per-triple vector and add it up in fixed order."""
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(e_triples, op=dist.ReduceOp.SUM)
    total = 0.0
    for x in e_triples.detach().cpu().tolist():
        total += x
    return total


def broadcast_inputs(tensors, dist=None, src=0):
    """Replicate rank `src`'s input tensors on every rank (NCCL broadcast over NVLink)."""
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
