"""Multi-GPU plan of the hot path: independent units, no exchange step.

Every output frame depends only on its own 7-frame window of LR frames + priors (test_LD_37.py:143-160) and per-frame
features depend only on that frame (arch/SIDECVSR_our.py:4417-4419), so the work list of (sequence, frame) items is
partitioned statically: one process per GPU, weights replicated, no collective on the data path.  The only collective
is the end-of-job gather of the metric sums (the reference computes PSNR/SSIM per sequence on the host,
metric/psnr_ssim.py:446-484): one all_reduce(SUM) of an [n_seq, 3] fp64 tensor (sum of per-frame PSNR, sum of per-frame
SSIM, frame count -- what metrics.psnr_ssim(accum=) accumulates), NCCL on the GPU box, gloo in the CPU tests.

Nothing here touches CUDA: the functions are plain host logic and are tested with world_size-2 gloo groups.
"""
import torch
import torch.distributed as dist

HALO = 3   # frames of temporal context on each side of an output frame (nframes // 2, arch:4377)


def partition(n_items, world, rank):
    """Contiguous block [lo, hi) of `n_items` owned by `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if world <= 0 or not 0 <= rank < world or n_items < 0:
        raise ValueError("partition: bad arguments")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sequence_shard(n_seq, world, rank):
    """Sequence ids owned by `rank` (config c4: 64 sequences over 2/4/8 GPUs -> blocks of 32/16/8)."""
    lo, hi = partition(n_seq, world, rank)
    return list(range(lo, hi))


def frame_shard(n_frames, world, rank):
    """Within one sequence (load balance of long clips): output frames [lo, hi) owned by `rank` and the frame range
    [read_lo, read_hi) it must read (a HALO-frame halo each side, clipped like generate_input_index, test_LD_37.py:13-16).
    The first window of a shard recomputes the per-frame features of its halo: the L1_fea cache is a pure optimisation."""
    lo, hi = partition(n_frames, world, rank)
    return (lo, hi), (max(lo - HALO, 0), min(hi + HALO, n_frames))


def window_indices(i, n_frames, radius=HALO):
    """Frame indices of the 7-frame window of output frame i, clipped at the sequence ends (test_LD_37.py:13-16)."""
    return [min(max(j, 0), n_frames - 1) for j in range(i - radius, i + radius + 1)]


def noise_key(seed, sequence, frame, neighbour):
    """Seed of the Gumbel uniforms of (sequence, frame, neighbour): a function of the work item, not of the device RNG state
    or of the rank that runs it -> outputs are identical for every sharding (same formula as synthetic.gumbel_uniforms)."""
    return 70000000 + ((seed * 4099 + sequence) * 4099 + frame) * 7 + neighbour


def gather_metrics(local, group=None):
    """Sum of the per-rank [n_seq, 3] fp64 metric sums over all ranks (in place; identity without a process group).
    Rows a rank does not own must be zero.  Returns the reduced tensor (see mean_metrics_per_sequence)."""
    if local.dtype != torch.float64 or local.dim() != 2 or local.size(1) != 3:
        raise ValueError("gather_metrics: expected an [n_seq, 3] float64 tensor")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local


def mean_metrics_per_sequence(sums):
    """(mean PSNR [dB], mean SSIM) of each sequence from the reduced [n_seq, 3] sums = (sum of per-frame PSNR, sum of per-frame
    SSIM, frames) that metrics.psnr_ssim(accum=) / FrameDriver.run accumulate: the per-sequence average of per-frame values, which is
    what the reference's cal_psnr_ssim reports (metric/psnr_ssim.py:477-481)."""
    n = sums[:, 2].clamp_min(1.0)
    return sums[:, 0] / n, sums[:, 1] / n