"""Gradient exchange of the data-parallel layerwise loop without a collective kernel (the reference has no
multi-GPU path of its own: base/base_trainer.py:16-19 wraps the model in a DataParallel that does not work for it, F11).

The student gradients (7.84 M fp32 for the 51M plan) become final site by site while the backward is still running.
`PeerGradBucket` keeps them in a SYMMETRIC buffer -- the same allocation on every rank, mapped into every peer's address
space over NVLink / NVSwitch (torch.distributed._symmetric_memory) -- laid out [parity][source rank][bucket]:

  * a rank writes its gradients into its own slot of its own buffer (the kernels' output pointers point there);
  * as soon as a region is final, `push(lo, hi)` copies it into the same slot of every PEER's buffer on a side stream: plain
    device-to-device copies, which the copy engines execute -- no SM is taken from the 148-CTA persistent kernels of the
    step (an NCCL kernel cannot co-reside with them, which is why the all-reduce used to be exposed at the end of the step);
  * `finish()` is one signal exchange on the side stream (every peer's pushes have landed) that the compute stream waits for;
  * the optimizer then reads the `world` copies it finds in LOCAL memory and steps on their mean, summed in rank order:
    `kdcc_radam_step_multi` -- the reduction is fused into the optimizer pass and is bit-identical on every rank.

Two parities alternate between steps so that a fast rank may already push step t+1 while a slow one still reads step t.
Falls back to nothing: if symmetric memory is unavailable the constructor raises and the caller keeps the NCCL all-reduce
(`GradBucket.all_reduce_mean`)."""
import torch
import torch.distributed as dist


class PeerGradBucket:
    def __init__(self, numel, device, group=None):
        import torch.distributed._symmetric_memory as symm
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGradBucket needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = torch.device(device)
        self.numel = int(numel)
        self.stride = -(-self.numel // 64) * 64                  # floats per slot: 256-byte aligned slots
        total = 2 * self.world * self.stride
        try:
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass
        self.buf = symm.empty(total, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, self.group)
        self.peers = {}
        for r in range(self.world):
            if r != self.rank:
                self.peers[r] = self.handle.get_buffer(r, (2, self.world, self.stride), torch.float32)
        self.mine = self.buf.view(2, self.world, self.stride)
        self.parity = 0
        self.comm = torch.cuda.Stream(device=self.device)
        self._ready = torch.cuda.Event()
        self._done = torch.cuda.Event()
        torch.cuda.synchronize(self.device)
        self.handle.barrier(channel=0)
        torch.cuda.synchronize(self.device)

    # ---- this step's gradient slot (what the kernels write / autograd accumulates into) ----
    def local(self):
        return self.mine[self.parity, self.rank, :self.numel]

    def sources(self):
        """(tensor view of source 0's copy, floats between sources, number of sources) for kdcc_radam_step_multi."""
        return self.mine[self.parity, 0, :self.numel], self.stride, self.world

    def push(self, lo=0, hi=None):
        """[lo, hi) of this step's gradients is final on the current stream: copy it to every peer on the side stream."""
        hi = self.numel if hi is None else hi
        if hi <= lo:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        src = self.mine[self.parity, self.rank, lo:hi]
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ev)
            for r, view in self.peers.items():
                view[self.parity, self.rank, lo:hi].copy_(src, non_blocking=True)

    def finish(self):
        """All pushes of this step, of every rank, have landed before anything later on the current stream runs."""
        with torch.cuda.stream(self.comm):
            self.handle.barrier(channel=0)
            self._done.record(self.comm)
        torch.cuda.current_stream(self.device).wait_event(self._done)

    def flip(self):
        self.parity ^= 1
