"""Channel sharding for multi-GPU runs: channels are independent sessions (reference: one dsp_worker thread per client,
src/dsp_worker.c:188), so GPU g of G owns a contiguous block of channels and there is no collective on the data path."""


def shard_channels(total_channels, world_size, rank):
    """Contiguous [first, first + count) owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(total_channels, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def channel_seed(job_seed, channel):
    """Input of channel c is a function of (job_seed, c) only, so any sharding sees the same signals."""
    return int(job_seed) * 1000003 + int(channel)
