"""Data-parallel plumbing for the learner (one process per GPU, torch.distributed).

Episodes are independent units: rank r trains on a contiguous slice of the sampled batch and
contributes the UN-normalised gradient of sum((td * mask)^2) plus the five loss sums, packed into one
fp32 buffer; ONE all-reduce(sum) per step, after which every rank applies the identical clip + RMSprop
update (parameters / optimizer state stay replicated - no broadcast).  Replay indices are drawn
once (rank 0's numpy stream) and sliced, so sampling stays identical to the 1-GPU run."""
import torch as th
import torch.distributed as dist


def is_active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_slice(batch_size, rank, world):
    """Contiguous episode range [lo, hi) of this rank (the first `batch_size % world` ranks get one more)."""
    base, rem = divmod(batch_size, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_fields(fields, rank, world):
    """Slice every batch-major field to this rank's episodes (views, no copy)."""
    some = next(iter(fields.values()))
    lo, hi = shard_slice(some.shape[0], rank, world)
    return {k: v[lo:hi] for k, v in fields.items()}


def rank():
    return dist.get_rank() if is_active() else 0


def world_size():
    return dist.get_world_size() if is_active() else 1


def allreduce_step(flat_grad_and_sums):
    """THE exchange of a step: sum `[un-normalised flat gradient | loss sums as (hi, lo) floats]` (one fp32 buffer, see
    pmb_dp_pack in include/pymarl_b200.h) over all ranks, in place, in ONE collective."""
    if not is_active():
        return
    dist.all_reduce(flat_grad_and_sums, op=dist.ReduceOp.SUM)


class PeerExchange:
    """The fused exchange + update over NVLink peer memory (csrc/dp_peer.cu, include/pymarl_b200.h): the rank's flat
    gradient lives in `buf` (a torch tensor the other ranks of the node map through CUDA IPC) and ONE kernel per rank does
    the ordered all-reduce, the grad norm, clip and RMSprop.  Built collectively (every rank must construct it at the same
    point); `ok` is False - on EVERY rank - when any rank could not export / map (expandable-segment allocator, no peer
    access, more than 8 ranks, ranks on different nodes), and the caller falls back to the NCCL all-reduce."""

    def __init__(self, n, device):
        import ctypes as C
        from . import _lib
        self.n, self.step, self.ok = n, 0, False
        self.world, self.rank = world_size(), rank()
        L = _lib.lib()
        self.buf = th.zeros(L.pmb_dp_exchange_floats(n), dtype=th.float32, device=device)
        self.scratch = th.zeros(4 * n + 4096, dtype=th.uint8, device=device)
        self._mapped = []
        handle, off, good = bytes(64), 0, 1
        try:
            if self.world > 8:
                raise RuntimeError("more than 8 ranks")
            hbuf = C.create_string_buffer(64)
            o = C.c_int64(0)
            _lib.check(L.pmb_ipc_export(_lib.ptr(self.buf), hbuf, C.byref(o)), "pmb_ipc_export")
            handle, off = hbuf.raw, o.value
        except Exception:
            good = 0
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (good, handle, off, _hostname()))
        good = int(all(g[0] for g in gathered) and len({g[3] for g in gathered}) == 1)
        ptrs = (C.c_void_p * self.world)()
        if good:
            try:
                for r, (_, h, o_r, _) in enumerate(gathered):
                    if r == self.rank:
                        ptrs[r] = self.buf.data_ptr()
                    else:
                        # an allocation (torch allocator segment) is mapped once per process and then reused: a second
                        # exchange buffer that lands in the same segment must not open the same handle again
                        base = _OPENED.get(h)
                        if base is None:
                            p = C.c_void_p()
                            _lib.check(L.pmb_ipc_open(h, 0, C.byref(p)), "pmb_ipc_open")
                            base = _OPENED[h] = p.value
                        ptrs[r] = base + o_r
            except Exception:
                good = 0
        flag = th.tensor([good], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)            # also: nobody maps a buffer before everybody exported it
        self.ok = bool(flag.item())
        self.ptrs = ptrs

    def grad_view(self):
        """[n + DP_TAIL_FLOATS] floats at the head of the exchange buffer: the learner's flat gradient lives here."""
        from . import _lib
        return self.buf[:self.n + _lib.DP_TAIL_FLOATS]

    def fused_update(self, flat_p, flat_sq, flat_target, do_sync, stats, lr, alpha, eps, clip, stream):
        from . import _lib
        self.step += 1
        _lib.check(_lib.lib().pmb_dp_fused_allreduce_update(self.world, self.rank, self.ptrs, self.n, self.step, _lib.ptr(flat_p),
                                                            _lib.ptr(flat_sq), _lib.ptr(flat_target), int(do_sync),
                                                            _lib.ptr(stats), lr, alpha, eps, clip, _lib.ptr(self.scratch),
                                                            self.scratch.numel(), stream), "pmb_dp_fused_allreduce_update")

    def error_word(self):
        """0, or which spin timed out (1: a peer flag, 2: the grid barrier); reads the device (synchronises)."""
        off = ((4 * self.n + 255) // 256) * 256 + 1024 + 16
        return int(self.scratch[off:off + 4].view(th.int32).item())

    def close(self):
        """Mappings of peer allocations are shared by all exchanges of the process and live until it exits."""
        self.ok = False


_OPENED = {}          # IPC handle (64 bytes) -> base address of the mapping in this process


def _hostname():
    import socket
    return socket.gethostname()
