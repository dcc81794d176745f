"""The one exchange step of the sharded path: every rank all-gathers its padded descriptors [B,K,128], LAFs
[B,K,2,3] and counts [B] once per step.  Host-side plumbing over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Each step's results are packed into one staging block [B, K*134 + 1] and gathered with ONE collective that runs on the
backend's own stream while the next step computes (double-buffered); a slot is waited for before it is reused and `drain()`
closes a timed region."""
import torch
import torch.distributed as dist


class DescriptorExchange:
    def __init__(self, world, B, K, device):
        self.world, self.B, self.K = world, B, K
        self.row = K * 134 + 1
        self.stage = [torch.empty(B, self.row, device=device) for _ in range(2)]
        self.gath = [torch.empty(world * B, self.row, device=device) for _ in range(2)]
        self.work, self.i = [None, None], 0

    def submit(self, lafs, desc, count):
        """Queue the all-gather of this step's (lafs [B,K,2,3], desc [B,K,128], count [B] int32); returns immediately."""
        B, K = self.B, self.K
        s = self.i & 1
        if self.work[s] is not None:
            self.work[s].wait()
        st = self.stage[s]
        st[:, :K * 128].copy_(desc.reshape(B, -1)); st[:, K * 128:K * 134].copy_(lafs.reshape(B, -1)); st[:, K * 134].copy_(count)
        self.work[s] = dist.all_gather_into_tensor(self.gath[s], st, async_op=True)
        self.i += 1

    def drain(self):
        """Wait (on the current stream) for every queued all-gather."""
        for s in range(2):
            if self.work[s] is not None:
                self.work[s].wait()
                self.work[s] = None

    def last(self):
        """(desc [world*B,K,128], lafs [world*B,K,2,3], count [world*B] int32) of the most recently submitted step (after drain)."""
        K, n = self.K, self.world * self.B
        g = self.gath[(self.i - 1) & 1]
        return g[:, :K * 128].reshape(n, K, 128), g[:, K * 128:K * 134].reshape(n, K, 2, 3), g[:, K * 134].round().int()
