"""Multi-GPU sharding of the block stream: blocks are independent, so
the host deals contiguous batches of blocks round-robin to the ranks (batch k -> rank
k mod G) and the single ticket FIFO restores order.  No data-path collective."""


def batch_ranges(nblocks, batch, world):
    """[(rank, first_block, count)] in stream order."""
    out = []
    k = 0
    first = 0
    while first < nblocks:
        cnt = min(batch, nblocks - first)
        out.append((k % world, first, cnt))
        first += cnt
        k += 1
    return out


def my_ranges(nblocks, batch, world, rank):
    return [(f, c) for r, f, c in batch_ranges(nblocks, batch, world) if r == rank]


def merge_in_order(nblocks, batch, world, per_rank_outputs):
    """per_rank_outputs[rank] = list of encoded batches in that rank's own order."""
    cursors = [0] * world
    out = []
    for r, _f, _c in batch_ranges(nblocks, batch, world):
        out.append(per_rank_outputs[r][cursors[r]])
        cursors[r] += 1
    return b"".join(out)
