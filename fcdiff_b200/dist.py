"""
Edge sharding over the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch; ``gloo``
in the CPU tests).  Rank r owns the contiguous edge range
``[r*chunkC, min((r+1)*chunkC, C))`` of ``b`` / ``bt`` (rows) for the
edge-parallel kernels K2, K3, K4, and the patient range
``[r*chunkU, min((r+1)*chunkU, U))`` for the region sweep K2b, which is
sequential over regions but independent across patients.  Exchanges per EM
iteration:

* all-gather of ``lq_F`` / ``q_F`` (24*C bytes each) after K2,
* all-gather of ``lq_R`` / ``q_R`` (16*N*U bytes each) after K2b,
* all-reduce (sum) of <= 6 doubles after K3a, per K3b evaluation and after K4.

The reference has no distributed code; the sharded fit must equal the
single-device fit up to the order of the fp64 partial sums.
"""
import torch
import torch.distributed as dist


def _ceil_div(a, b):
    return (a + b - 1) // b


class EdgeShards(object):
    """Shard map + the three collectives of the sharded fit."""

    def __init__(self, group=None, rank=None, world=None):
        self.group = group
        if rank is None:
            rank = dist.get_rank(group)
        if world is None:
            world = dist.get_world_size(group)
        self.rank = int(rank)
        self.world = int(world)

    def key(self):
        return (self.rank, self.world)

    @staticmethod
    def chunk(total, world):
        return _ceil_div(total, world)

    def span(self, total, rank=None):
        """(start, length) of this rank's contiguous share of ``total`` items."""
        r = self.rank if rank is None else rank
        ch = self.chunk(total, self.world)
        start = min(r * ch, total)
        return start, min(ch, total - start)

    def ranges(self, C, U):
        (c0, Cl) = self.span(C)
        (u0, Ul) = self.span(U)
        return (c0, Cl, u0, Ul)

    # ------------------------------------------------------------------ collectives
    def allreduce_terms(self, out, idxs):
        """Sums the device vector ``out`` over ranks in place (one collective, no
        staging).  Entries in ``idxs`` are edge-local partial sums; the other
        entries are already complete on every rank, so their sum is divided by
        the world size by ``fix_replicated`` after the download."""
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def any_rank(self, flag):
        """True on every rank if ``flag`` is true on any rank (one tiny all-reduce):
        used where the ranks must agree on a code path whose collectives differ."""
        dev = "cuda" if (torch.cuda.is_available() and dist.get_backend(self.group) == "nccl") else "cpu"
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return bool(t.item() > 0)

    def fix_replicated(self, host_vec, idxs):
        """Undo the summation of entries that were complete on every rank."""
        keep = set(idxs)
        for i in range(len(host_vec)):
            if i not in keep:
                host_vec[i] = host_vec[i] / self.world
        return host_vec

    def allgather_edges(self, lqF, qF, C):
        """lqF, qF: flat [C*3]; this rank's edge rows are valid on entry, all rows on
        exit.  One collective for both arrays."""
        ch = self.chunk(C, self.world)
        (start, length) = self.span(C)
        local = lqF.new_zeros((2, ch * 3))
        local[0, :length * 3].copy_(lqF[start * 3:(start + length) * 3])
        local[1, :length * 3].copy_(qF[start * 3:(start + length) * 3])
        gathered = lqF.new_empty((self.world, 2, ch * 3))
        dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        lqF.copy_(gathered[:, 0].reshape(-1)[:C * 3])
        qF.copy_(gathered[:, 1].reshape(-1)[:C * 3])

    def allgather_patients(self, lqR, qR, N, U):
        """lqR, qR: flat [N*U*2]; this rank's patient columns are valid on entry, all
        columns on exit.  One collective for both arrays."""
        ch = self.chunk(U, self.world)
        (u0, Ul) = self.span(U)
        local = lqR.new_zeros((2, N, ch, 2))
        local[0, :, :Ul].copy_(lqR.view(N, U, 2)[:, u0:u0 + Ul])
        local[1, :, :Ul].copy_(qR.view(N, U, 2)[:, u0:u0 + Ul])
        gathered = lqR.new_empty((self.world, 2, N, ch, 2))
        dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        for (i, t) in enumerate((lqR, qR)):
            t.view(N, U, 2).copy_(gathered[:, i].permute(1, 0, 2, 3).reshape(N, self.world * ch, 2)[:, :U])

    def exchange_patient_blocks(self, bt_local, Cl, C, U):
        """Edge-sharded -> patient-sharded re-layout of the patient correlations
        (one all-to-all at set-up): rank r holds rows [c0, c0+Cl) of bt for all
        patients and receives every edge of its own patients.  Returns the
        contiguous (C, Ul) block."""
        (u0, Ul) = self.span(U)
        send, in_split, out_split = [], [], []
        for r in range(self.world):
            (ur, Ur) = self.span(U, r)
            send.append(bt_local[:, ur:ur + Ur].reshape(-1))
            in_split.append(Cl * Ur)
            (_, Cr) = self.span(C, r)
            out_split.append(Cr * Ul)
        sendbuf = torch.cat(send) if send else bt_local.new_empty((0,))
        recv = bt_local.new_empty((C * Ul,))
        dist.all_to_all_single(recv, sendbuf, output_split_sizes=out_split, input_split_sizes=in_split,
                               group=self.group)
        return recv.view(C, Ul)


def init_from_env(backend="nccl"):
    """Initialises ``torch.distributed`` from torchrun's environment, binds the
    local GPU and returns an :class:`EdgeShards` (None when WORLD_SIZE is 1)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return EdgeShards()
