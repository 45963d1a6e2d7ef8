"""Multi-GPU plumbing: one process per GPU (torchrun), ``torch.distributed`` only.

* Inference shards genes (or (checkpoint, gene-chunk) units) across ranks — no data-path
  collective; ``gather_sharded`` merely collects the per-rank logits at the end.
* Training is data parallel with exactly one exchange per step: a SUM all-reduce of the
  gradient-receiving prefix of the flat gradient buffer (one bucket); the 1/world mean is
  folded into the fused AdamW launch.
Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU tensors (unit tests).
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise from the torchrun environment; returns (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU BEFORE it allocates pinned host buffers, so that
    the staging memory of every rank sits on the NUMA node (and behind the PCIe root) of that rank's GPU instead of
    wherever the launcher happened to start the process.  Returns a short description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < n_cpu]
        if not cpus:
            return "nvml reported no local CPUs"
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "local CPUs outside this process's cpuset"
        os.sched_setaffinity(0, allowed)
        return f"{len(allowed)} CPUs local to GPU {local_rank} ({allowed[0]}..{allowed[-1]})"
    except Exception as e:                                          # noqa: BLE001 - best effort, never fatal
        return f"unavailable ({type(e).__name__})"


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of `n_items` units for `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sweep_units(n_checkpoints, n_genes, chunk):
    """(checkpoint, gene_lo, gene_hi) work units of an ensemble sweep (BASELINE configs[4])."""
    return [(c, lo, min(n_genes, lo + chunk)) for c in range(n_checkpoints) for lo in range(0, n_genes, chunk)]


def allreduce_gradients(flat_grad, n_active, group=None):
    """SUM all-reduce of the single gradient bucket; returns the scale (1/world) still to apply."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(flat_grad[:n_active], op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def gather_sharded(local, n_total, group=None):
    """Concatenate per-rank results of a shard_range() split on every rank (host-side gather)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    base = (n_total + world - 1) // world
    pad = torch.zeros((base,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(n_total, r, world)
        out.append(p[:hi - lo])
    return torch.cat(out)
