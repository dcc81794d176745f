"""The one exchange step of the sharded path: every rank all-gathers its padded descriptors [B,K,128], LAFs
[B,K,2,3] and counts [B] once per step.  Host-side plumbing over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

A rank's results of one step are ONE flat float32 block  [desc B*K*128 | lafs B*K*6 | count B (int32 bits)]  that the pipeline's
kernels write directly (`DetectDescribePipeline(..., outputs=exchange.outputs())`: the HardNet head, the LAF writer and the count
writer store into views of the block), so a step costs exactly one collective and no packing copies.  There are two blocks: the
gather of step i runs on the backend's own stream while step i+1 computes into the other block; a block is waited for before it is
reused and `drain()` closes a timed region."""
import torch
import torch.distributed as dist


class DescriptorExchange:
    def __init__(self, world, B, K, device, n_slots=2):
        self.world, self.B, self.K, self.n_slots = world, B, K, n_slots
        self.block = B * K * 134 + B
        self.stage = [torch.zeros(self.block, dtype=torch.float32, device=device) for _ in range(n_slots)]
        self.gath = [torch.empty(world * self.block, dtype=torch.float32, device=device) for _ in range(n_slots)]
        self.work, self.i = [None] * n_slots, 0

    def views(self, flat):
        """(lafs [B,K,2,3], desc [B,K,128], count [B] int32) views of one rank's block."""
        B, K = self.B, self.K
        return (flat[B * K * 128:B * K * 134].view(B, K, 2, 3), flat[:B * K * 128].view(B, K, 128), flat[B * K * 134:].view(torch.int32))

    def outputs(self):
        """Per slot: the tensors a producer writes its step results into (pass to DetectDescribePipeline(outputs=...))."""
        return [self.views(s) for s in self.stage]

    def submit(self, slot=None):
        """Queue the all-gather of block `slot` (default: round robin); returns immediately.  The producer must have ENQUEUED its writes
        into the block on the current stream before this call, and must not enqueue writes into the same block again before the
        next submit() of that slot (which waits for the gather)."""
        s = (self.i % self.n_slots) if slot is None else slot
        if self.work[s] is not None:
            self.work[s].wait()
        self.work[s] = dist.all_gather_into_tensor(self.gath[s], self.stage[s], async_op=True)
        self.last_slot = s
        self.i += 1

    def wait_slot(self, slot):
        """Before a producer overwrites block `slot`: its previous gather must have read it."""
        if self.work[slot] is not None:
            self.work[slot].wait()
            self.work[slot] = None

    def drain(self):
        """Wait (on the current stream) for every queued all-gather."""
        for s in range(self.n_slots):
            if self.work[s] is not None:
                self.work[s].wait()
                self.work[s] = None

    def last(self):
        """(desc [world*B,K,128], lafs [world*B,K,2,3], count [world*B] int32) of the most recently submitted step (after drain)."""
        B, K, W = self.B, self.K, self.world
        g = self.gath[self.last_slot].view(W, self.block)
        return (g[:, :B * K * 128].reshape(W * B, K, 128), g[:, B * K * 128:B * K * 134].reshape(W * B, K, 2, 3),
                g[:, B * K * 134:].contiguous().view(torch.int32).reshape(W * B))


class CopyEngineExchange(DescriptorExchange):
    """The same exchange without a collective kernel: the gather buffers are symmetric memory (torch.distributed._symmetric_memory: peer-
    mapped over NVLink), every rank pushes its block into every peer's buffer with peer-to-peer cudaMemcpyAsync (the DMA copy engines move
    it; no SM is taken from the step that computes meanwhile), and a symmetric-memory barrier marks the step complete.  Raises at
    construction when symmetric memory is unavailable (callers fall back to DescriptorExchange)."""

    def __init__(self, world, B, K, device, n_slots=2):
        import torch.distributed._symmetric_memory as symm
        self.world, self.B, self.K, self.n_slots = world, B, K, n_slots
        self.rank = dist.get_rank()
        self.block = B * K * 134 + B
        self.stage = [torch.zeros(self.block, dtype=torch.float32, device=device) for _ in range(n_slots)]
        self.gath = [symm.empty(world * self.block, dtype=torch.float32, device=device) for _ in range(n_slots)]
        self.hdl = [symm.rendezvous(g, dist.group.WORLD) for g in self.gath]
        self.peer = [[h.get_buffer(p, (world * self.block,), torch.float32) for p in range(world)] for h in self.hdl]
        self.stream = torch.cuda.Stream(device=device)
        self.done = [None] * n_slots
        self.i = 0

    def submit(self, slot=None):
        s = (self.i % self.n_slots) if slot is None else slot
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.stream.wait_event(ev)                       # the step's kernels have written the block
        lo = self.rank * self.block
        with torch.cuda.stream(self.stream):
            for p in range(self.world):
                self.peer[s][p][lo:lo + self.block].copy_(self.stage[s], non_blocking=True)
            self.hdl[s].barrier(channel=s)               # every rank's pushes into this slot have landed
            d = torch.cuda.Event()
            d.record(self.stream)
        self.done[s] = d
        self.last_slot = s
        self.i += 1

    def wait_slot(self, slot):
        if self.done[slot] is not None:
            torch.cuda.current_stream().wait_event(self.done[slot])
            self.done[slot] = None

    def drain(self):
        for s in range(self.n_slots):
            self.wait_slot(s)


def make_exchange(world, B, K, device, kind=None):
    """kind: "ce" (copy engines over symmetric memory), "nccl" (all_gather_into_tensor) or None = $AG_EXCHANGE, default "ce" with a
    fallback to "nccl" when symmetric memory cannot be set up.  Returns (exchange, kind actually used)."""
    import os
    kind = kind or os.environ.get("AG_EXCHANGE", "ce")
    if kind == "ce" and device.type == "cuda":
        try:
            return CopyEngineExchange(world, B, K, device), "ce"
        except Exception as e:   # noqa: BLE001
            import sys
            print("affnet_b200.exchange: symmetric memory unavailable (%s): using the NCCL all-gather" % str(e)[:200], file=sys.stderr)
    return DescriptorExchange(world, B, K, device), "nccl"
