"""K-mer sharding across GPUs and the gather of the per-variant result table.

Variants are independent given the once-per-run state, so the multi-GPU plan is a contiguous
range of variant ids per rank and one collective on the way out: the gather of the result
table on rank 0, ordered by rank so that output order == input order (the reference keeps
input order through ``pool.starmap``, __main__.py:541, :777).  ``torch.distributed`` is the
plumbing only (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import numpy as np

#: result-table columns gathered across ranks: (name, bytes per variant, numpy dtype)
TABLE_COLUMNS = (('carriers', 4, np.int32), ('missing', 4, np.int32), ('af', 8, np.float64),
                 ('prep', 8, np.float64), ('pvalue', 8, np.float64), ('beta', 8, np.float64),
                 ('bse', 8, np.float64), ('extra', 8, np.float64), ('flags', 4, np.uint32))
ROW_BYTES = sum(b for _, b, _ in TABLE_COLUMNS)


def shard_range(n_variants, rank, world):
    """Contiguous [first, last) range of variant ids of ``rank``; sizes differ by at most 1."""
    base, rem = divmod(int(n_variants), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def column_offsets(n_rows):
    """Byte offset of every column in a packed structure-of-arrays table of ``n_rows``."""
    off, out = 0, {}
    for name, b, _ in TABLE_COLUMNS:
        out[name] = off
        off += n_rows * b
    return out


def table_pointers(base_ptr, n_rows):
    """Column name -> raw address inside a packed table buffer (for Engine.fetch_into)."""
    return {name: base_ptr + off for name, off in column_offsets(n_rows).items()}


def unpack_table(buf, n_rows):
    """Packed table bytes (numpy uint8) -> dict of column arrays (views)."""
    out = {}
    offs = column_offsets(n_rows)
    for name, b, dt in TABLE_COLUMNS:
        out[name] = buf[offs[name]:offs[name] + n_rows * b].view(dt)
    return out


def gather_tables(table, n_rows_per_rank, dst=0):
    """Gather every rank's packed table (a torch uint8 tensor, CPU for gloo or CUDA for nccl)
    on ``dst``.  Returns the list of per-rank tensors on ``dst`` (in rank order), else None.
    Shards may differ in length by one row: buffers are padded to the longest."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    rank = dist.get_rank()
    longest = max(n_rows_per_rank) * ROW_BYTES
    if table.numel() < longest:
        table = torch.cat([table, torch.zeros(longest - table.numel(), dtype=torch.uint8,
                                              device=table.device)])
    bufs = [torch.empty_like(table) for _ in range(world)] if rank == dst else None
    dist.gather(table, bufs, dst=dst)
    if rank != dst:
        return None
    return [b[:n * ROW_BYTES] for b, n in zip(bufs, n_rows_per_rank)]


def merge_tables(per_rank, n_rows_per_rank):
    """Per-rank packed tables (numpy uint8) -> dict of full-length columns in variant order."""
    cols = {name: [] for name, _, _ in TABLE_COLUMNS}
    for buf, n in zip(per_rank, n_rows_per_rank):
        for name, arr in unpack_table(buf, n).items():
            cols[name].append(arr)
    return {name: np.concatenate(v) for name, v in cols.items()}
