"""Independent Markov chains, one per GPU, and the pooled statistics that combine them.

The path does not shard ("replicas only"): every rank runs its own chain with seed
`SEED + rank` and nothing crosses NVLink during sweeps.  After measuring, per-chain bins are combined
exactly like the reference's offline pooling `combined_mean_and_var(ns, mus, vs)` (src/statistics.jl:22-36):
one all-reduce (SUM) of the packed moments (n, n*mu, (n-1)*v + n*|mu|^2) per observable.
"""
import numpy as np
import torch
import torch.distributed as dist


def init_process_group(backend=None, device_id=None):
    """Rendezvous from the torchrun environment (RANK / WORLD_SIZE / MASTER_*); no-op for a single process."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if device_id is not None:
        dist.init_process_group(backend=backend, device_id=device_id)
    else:
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def combined_mean_and_var(n, mean, var, device=None):
    """Pooled mean and (unbiased) variance over all ranks of per-rank samples of length n.

    `mean`/`var` may be real or complex arrays (variance is real).  Reproduces statistics.jl:22-36:
        meanc = sum_k n_k mu_k / nsum
        varc  = sum_k [(n_k - 1) v_k + n_k |mu_k - meanc|^2] / (nsum - 1)
    using sum_k n_k |mu_k - meanc|^2 = sum_k n_k |mu_k|^2 - nsum |meanc|^2, so a single all-reduce suffices.
    """
    mean = np.asarray(mean)
    var = np.asarray(var, dtype=np.float64)
    is_c = np.iscomplexobj(mean)
    m = mean.astype(np.complex128).ravel()
    packed = np.concatenate([[float(n)], n * m.real, n * m.imag, ((n - 1) * var + n * np.abs(mean) ** 2).ravel()])
    t = torch.from_numpy(packed)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    a = t.numpy()
    k = m.size
    nsum = a[0]
    meanc = (a[1:1 + k] + 1j * a[1 + k:1 + 2 * k]) / nsum
    varc = (a[1 + 2 * k:] - nsum * np.abs(meanc) ** 2) / (nsum - 1)
    meanc = meanc.reshape(mean.shape)
    return (meanc if is_c else meanc.real), varc.reshape(var.shape)
