"""Multi-GPU: batch-sharded sampling, one process per GPU, one all-gather of the finished images.

The path has no exchange step (no op mixes samples; the reference fans out N independent processes,
scripts/sample_diffusion.py:88-100, tools/frido/eval_layout2i_multiGPU.sh:9-12), so the only collective is
the gather of decoded images at the end of a batch."""
import torch
import torch.distributed as dist


def shard_bounds(n_global, rank, world):
    """Contiguous shard [lo, hi) of `n_global` samples for `rank`; remainders go to the first ranks."""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank=None, world=None):
    """Slice a globally-generated tensor (context, start noise) so results do not depend on the GPU count."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi].contiguous()


def gather_images(img, n_global=None, group=None):
    """All-gather per-rank image batches [b_r, C, H, W] into [n_global, C, H, W] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return img
    world = dist.get_world_size(group)
    n_global = img.shape[0] * world if n_global is None else n_global
    sizes = [shard_bounds(n_global, r, world) for r in range(world)]
    if all(hi - lo == img.shape[0] for lo, hi in sizes):
        out = torch.empty((world * img.shape[0],) + tuple(img.shape[1:]), dtype=img.dtype, device=img.device)
        dist.all_gather_into_tensor(out, img.contiguous(), group=group)
        return out
    # ragged shards: collectives need equal sizes -> pad to the largest shard, gather, trim
    mx = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((mx,) + tuple(img.shape[1:]), dtype=img.dtype, device=img.device)
    padded[: img.shape[0]] = img
    out = torch.empty((world * mx,) + tuple(img.shape[1:]), dtype=img.dtype, device=img.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


@torch.no_grad()
def sample_sharded(model, sampler, S, shape, context_global, init_noise_global, num_stage, **kw):
    """Every rank samples its shard and decodes it; returns the gathered images of the whole batch."""
    ctx = shard(context_global).to(model.device)
    x0 = shard(init_noise_global).to(model.device)
    z, _ = sampler.sample(S, ctx.shape[0], shape, conditioning=ctx, num_stage=num_stage, init_noise=x0, verbose=False, **kw)
    img = model.decode_first_stage(z)
    return gather_images(img, n_global=context_global.shape[0])
