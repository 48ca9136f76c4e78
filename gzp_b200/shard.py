"""Multi-GPU sharding of the block stream: blocks are independent, so
the host deals contiguous batches of blocks round-robin to the GPUs (batch k -> GPU
k mod G) and one ordered drain restores the stream.  No data-path collective.
This is the dealing rule of gzpb_encode_stream_multi / gzpb_writer_create_multi (gzpb_api.cu) stated in
Python, so that the world_size-2 gloo test (tests/test_shard_gloo.py) can hold it on CPU."""


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
