"""Multi-GPU plumbing: particles / edges are independent (SimulatorInterface::ForwardSimulateRobots contract,
simple_simulator_interface.hpp:623; clustering is per edge), so the batch is sharded in contiguous ranges with
the SDF and robot tables replicated, and the only exchange is ONE gather of the outcome clusters
(label, collided flag, configuration per particle) at the end of a step.  torch.distributed is the transport:
NCCL over NVLink on the GPUs, gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous, balanced [lo, hi) range of `total` items owned by `rank`."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def record_bytes(width):
    """Bytes per particle of an outcome record: int32 label + uint8 collided (padded to 4) + `width` doubles."""
    return 4 + 4 + 8 * width


def pack_outcomes(labels, collided, configs_soa):
    """labels int32 [n], collided uint8 [n], configs_soa float64 [C, n] -> one uint8 buffer (fixed-stride planes)."""
    n = labels.shape[0]
    coll32 = collided.to(torch.int32)
    return torch.cat([labels.contiguous().view(torch.uint8).reshape(-1), coll32.contiguous().view(torch.uint8).reshape(-1),
                      configs_soa.contiguous().view(torch.uint8).reshape(-1)]), n


def unpack_outcomes(buf, n, width):
    labels = buf[: 4 * n].view(torch.int32)
    collided = buf[4 * n: 8 * n].view(torch.int32).to(torch.uint8)
    configs = buf[8 * n: 8 * n + 8 * width * n].view(torch.float64).reshape(width, n)
    return labels, collided, configs


def gather_outcomes(packed, world, async_op=False):
    """The single collective of a step: every rank receives every rank's outcome records (equal shard sizes).
    async_op=True returns (out, work): the collective runs on the process group's own stream and the caller's stream
    does not wait for it until work.wait() - the next batch's simulation overlaps the exchange."""
    if world == 1:
        return (packed.unsqueeze(0), None) if async_op else packed.unsqueeze(0)
    out = torch.empty((world, packed.numel()), dtype=torch.uint8, device=packed.device)
    try:
        work = dist.all_gather_into_tensor(out, packed, async_op=async_op)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(parts, packed)
        out = torch.stack(parts)
        work = None
    return (out, work) if async_op else out


def merge_cluster_counts(gathered, n, width, world):
    """Host-side view of the gathered outcomes: per rank (labels, collided, configs) as numpy arrays."""
    res = []
    for r in range(world):
        l, c, q = unpack_outcomes(gathered[r], n, width)
        res.append((l.cpu().numpy(), c.cpu().numpy().astype(bool), np.ascontiguousarray(q.cpu().numpy().T)))
    return res
