"""Scene-sharded data parallelism for the backbone + the one collective of the north star.

The hot path shards by scene: every op indexes a batch element independently (the reference kernels
use ``blockIdx.x = batch``, e.g. lib/pointnet2/_ext_src/src/sampling_gpu.cu:82), BatchNorm statistics
are per replica in the reference (torch.nn.DataParallel, scripts/joint_scripts/train_3dvlp.py:124-126),
so one process per GPU runs its contiguous chunk of scenes with NO data-path collective.

The only exchange is the all-gather of L2-normalised per-object / per-sentence embeddings for the
OCC/OSC contrastive losses, modelled on the reference's (never-called) helper
``concat_all_gather`` (utils/utils_fn.py:10-21: no-grad ``all_gather`` + ``cat(dim=0)``).
Backend: NCCL over NVLink/NVSwitch on GPUs, gloo on CPU (tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous [start, end) chunk of `total` scenes owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_scenes(batch, world_size=None, rank=None):
    """Slice dim 0 of a tensor (or every tensor of a dict) down to this rank's scenes."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    if isinstance(batch, dict):
        return {k: shard_scenes(v, world_size, rank) if torch.is_tensor(v) else v for k, v in batch.items()}
    s, e = shard_range(batch.shape[0], world_size, rank)
    return batch[s:e]


@torch.no_grad()
def concat_all_gather(tensor):
    """utils/utils_fn.py:10-21: gather `tensor` from every rank and concatenate on dim 0 (no grad).
    Ranks must hold equally shaped tensors."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    out = [torch.empty_like(tensor) for _ in range(dist.get_world_size())]
    dist.all_gather(out, tensor.contiguous(), async_op=False)
    return torch.cat(out, dim=0)


def all_gather_with_local_grad(tensor):
    """Contrastive-loss gather: every rank sees all embeddings as negatives, and the slice that
    belongs to this rank keeps its autograd history (the MoCo/ALBEF pattern the reference helper
    was taken from)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    gathered = [torch.empty_like(tensor) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, tensor.detach().contiguous())
    gathered[dist.get_rank()] = tensor
    return torch.cat(gathered, dim=0)


def gather_contrastive_embeddings(obj_emb, sent_emb, normalize=True):
    """L2-normalise (as models/jointnet/jointnet.py:148-149 does for vote features) and all-gather
    the (B_loc*256, 128) object and (B_loc*lang_num_max, 128) sentence embeddings of SURVEY.md 8(e).
    Both tensors go out in ONE collective (a single flattened buffer): the exchange is
    latency-bound (~1 MB per GPU), so one launch beats two."""
    if normalize:
        obj_emb = torch.nn.functional.normalize(obj_emb, dim=-1)
        sent_emb = torch.nn.functional.normalize(sent_emb, dim=-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return obj_emb, sent_emb
    no, ns_ = obj_emb.shape[0], sent_emb.shape[0]
    flat = torch.cat([obj_emb.reshape(-1), sent_emb.reshape(-1)])
    world = dist.get_world_size()
    out = torch.empty(world * flat.numel(), device=flat.device, dtype=flat.dtype)
    dist.all_gather_into_tensor(out, flat.detach().contiguous())
    out = out.view(world, flat.numel())
    d = obj_emb.shape[-1]
    objs = [out[r, : no * d].view(no, d) for r in range(world)]
    sents = [out[r, no * d:].view(ns_, d) for r in range(world)]
    objs[dist.get_rank()] = obj_emb
    sents[dist.get_rank()] = sent_emb
    return torch.cat(objs, dim=0), torch.cat(sents, dim=0)
