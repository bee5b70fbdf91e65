"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch on the
GPU box, gloo in the CPU tests).  The hot path has exactly one collective per learn_batch: an
all-reduce(sum) of the flat gradient vector (with the step scalars riding at its tail)."""
import os

import numpy as np
import torch
import torch.distributed as td


_local = {"depth": 0}


class local_mode:
    """`with dist.local_mode(): ...` -- inside a multi-rank job, build and run networks that are NOT data-parallel
    (world_size() reports 1): single-rank baselines in parity tests, per-rank evaluation."""

    def __enter__(self):
        _local["depth"] += 1
        return self

    def __exit__(self, *exc):
        _local["depth"] -= 1


def initialized():
    return td.is_available() and td.is_initialized() and _local["depth"] == 0


def world_size():
    return td.get_world_size() if initialized() else 1


def rank():
    return td.get_rank() if initialized() else 0


def init_from_env(backend=None):
    """Initialise from the torchrun environment (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT)."""
    if initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        backend = backend or "nccl"
    else:
        backend = backend or "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    td.init_process_group(backend=backend)


def allreduce_sum(flat):
    """In-place sum over ranks of a DeviceArray (or torch tensor)."""
    if world_size() == 1:
        return flat
    t = flat if isinstance(flat, torch.Tensor) else flat.t
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return flat


class SymmetricBuffer:
    """The receive side of the fused all-reduce + optimizer kernel (`bf_opt_step_allreduce`, push form): a float32
    buffer of 2 (step parity) x world (source rank) x n slots in NVLink-mapped symmetric memory that every peer
    stores its gradient vector into.  `torch.distributed._symmetric_memory` does the allocation and the handle
    exchange (plumbing); the kernel that uses the peer pointers is ours."""
    BLOCKS = 64

    def __init__(self, tensor, handle, n):
        self.t = tensor
        self.handle = handle
        self.n = n
        self.rank, self.world = handle.rank, handle.world_size
        self.peers_dev = int(handle.buffer_ptrs_dev)
        self.pads_dev = int(handle.signal_pad_ptrs_dev)
        cap = int(handle.signal_pad_size) // (4 * self.world)
        self.blocks = max(1, min(self.BLOCKS, cap))
        self.epoch = torch.zeros(self.blocks, dtype=torch.int32, device=tensor.device)
        # barrier time-out word: mapped host memory, written by the kernel, read by the host for free
        self.timeout = torch.zeros(4, dtype=torch.int32).pin_memory()

    def timed_out(self):
        """True if a peer failed to reach the exchange of some step (valid once that step has been synchronised,
        e.g. after its scalars were read back)."""
        return bool(int(self.timeout[0]))

    def check(self):
        if self.timed_out():
            raise RuntimeError("data-parallel step failed: a peer rank did not reach the gradient exchange within the "
                               "time-out (crashed rank or mismatched learn_batch sequence); this replica was left un-updated")


_symm_failed = False


def symmetric_recv(n):
    """Collective (all ranks, same order): the symmetric receive buffer for gradient vectors of n floats, or None when
    the job is single-rank, not on NCCL, switched off (BF_FUSED_ALLREDUCE=0) or the platform has no peer mapping --
    the caller then keeps the NCCL all-reduce."""
    global _symm_failed
    if world_size() == 1 or _symm_failed or td.get_backend() != "nccl" or os.environ.get("BF_FUSED_ALLREDUCE", "1") == "0":
        return None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(2 * world_size() * n, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        handle = symm_mem.rendezvous(t, td.group.WORLD)
        t.zero_()
        torch.cuda.synchronize()
        td.barrier()
        return SymmetricBuffer(t, handle, n)
    except Exception as exc:   # noqa: BLE001  (any failure here only costs the fast path)
        _symm_failed = True
        if rank() == 0:
            print("[brainforge_b200] symmetric memory unavailable (%s: %s); using the NCCL all-reduce"
                  % (type(exc).__name__, exc), flush=True)
        return None


def global_count(m_local):
    """Sum of per-rank batch sizes, assuming equal shards (m_local * world_size: no collective needed).  The step
    itself carries every rank's true count in the scalar tail of the gradient vector, so `Backpropagation` detects
    an uneven split when it reads the step's scalars back and raises instead of training with a wrong 1/m."""
    return m_local * world_size()


def shard_bounds(n, r=None, ws=None):
    """Contiguous, even split of n units over ranks; the first n % ws ranks get one extra."""
    r = rank() if r is None else r
    ws = world_size() if ws is None else ws
    base, rem = divmod(n, ws)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def gather_concat(local, counts):
    """All-gather of variable-length 1-D float arrays (fitness values) -> concatenated host array."""
    if world_size() == 1:
        return np.asarray(local)
    dev = "cuda" if td.get_backend() == "nccl" else "cpu"
    mx = max(counts)
    buf = torch.zeros(mx, dtype=torch.float64, device=dev)
    buf[:len(local)] = torch.as_tensor(np.asarray(local, dtype=np.float64), device=dev)
    outs = [torch.zeros(mx, dtype=torch.float64, device=dev) for _ in counts]
    td.all_gather(outs, buf)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(outs, counts)])
