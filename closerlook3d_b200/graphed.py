"""CUDA-graph replay of one LocalAggregation training step (forward + backward).

The hot path of a small cloud batch is ~40 short kernels; launched eagerly from Python the host cannot keep
the GPU busy (measured on B200, BASELINE configs[1]: 1.3 ms of host time per step for 0.5 ms of kernels).
`GraphedStep` captures the whole step -- neighbour search, layout change, fused aggregation, BN, and the
autograd backward -- once into a CUDA graph with static input / output buffers and replays it.  Nothing is
cached between replays: every replay runs the full search and aggregation on whatever the static input
buffers hold at that moment.
"""
import torch


class GraphedStep:
    """step = module(xyz, xyz, mask, mask, features) ; out.backward(grad_out).

    static buffers (write your batch into them, then call replay()):
        .xyz (B,N,3) f32   .mask (B,N) i32   .features (B,C,N) f32   .grad_out (B,C_out,N) f32
    results after replay(): .out (B,C_out,N), .features.grad, parameter .grad tensors (static as well).
    """

    def __init__(self, module, xyz, mask, features, grad_out, warmup=3):
        from . import pt_utils
        self.module = module
        self.xyz = xyz.clone()
        self.mask = mask.clone()
        self.features = features.detach().clone().requires_grad_(True)
        self.grad_out = grad_out.clone()
        self.params = [p for p in module.parameters() if p.requires_grad]
        self._cache_was = pt_utils.cache_enabled
        pt_utils.cache_enabled = False  # a captured search must never be skipped on replay
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        try:  # parameters were created on the default stream; the captured backward accumulates on the capture stream
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except Exception:
            pass
        self.graph = torch.cuda.CUDAGraph()
        self.features.grad = None
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(self.graph):
            self.out = self.module(self.xyz, self.xyz, self.mask, self.mask, self.features)
            self.out.backward(self.grad_out)
        pt_utils.cache_enabled = self._cache_was

    def _eager(self):
        self.features.grad = None
        for p in self.params:
            p.grad = None
        out = self.module(self.xyz, self.xyz, self.mask, self.mask, self.features)
        out.backward(self.grad_out)
        return out

    def load(self, xyz, mask, features, non_blocking=True):
        """copy a batch (device or pinned host tensors) into the static input buffers"""
        self.xyz.copy_(xyz, non_blocking=non_blocking)
        self.mask.copy_(mask, non_blocking=non_blocking)
        with torch.no_grad():
            self.features.copy_(features, non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.out


class PipelinedTrainer:
    """End-to-end stepping from HOST batches: two GraphedSteps (double-buffered static inputs) and a copy stream,
    so the host->device copy of batch i+1 overlaps the replay of batch i; the step's result (sum of the output,
    a stand-in for the loss) is copied back asynchronously and read one step later.  Every batch is still
    copied, searched, aggregated and differentiated in full -- nothing is cached or skipped.

        tr = PipelinedTrainer(module, example_xyz, example_mask, example_features, grad_out)
        for host_batch in loader:            # pinned host tensors
            prev_loss = tr.step(host_batch)  # returns the result of the PREVIOUS step (None on the first)
        last = tr.flush()
    """

    def __init__(self, module, xyz, mask, features, grad_out, after_step=None):
        self.slots = [GraphedStep(module, xyz, mask, features, grad_out) for _ in range(2)]
        # high priority: its own hardware queue, so the H2D copies are never ordered behind the replay's kernels
        self.copy = torch.cuda.Stream(priority=-1)
        self.after_step = after_step            # e.g. the gradient all-reduce + optimizer step
        self.result_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]     # replay i finished (its buffers are free again)
        self.loaded = [torch.cuda.Event(), torch.cuda.Event()]   # inputs of slot i are on the device
        self.i = 0
        self.pending = None
        for e in self.done:
            e.record()

    def _issue_copy(self, slot, batch):
        s = self.slots[slot]
        self.copy.wait_event(self.done[slot])   # the previous replay on this slot no longer reads its inputs
        with torch.cuda.stream(self.copy):
            s.load(batch["xyz"], batch["mask"], batch["features"], non_blocking=True)
            self.loaded[slot].record(self.copy)

    def step(self, batch):
        slot = self.i & 1
        cur = torch.cuda.current_stream()
        self._issue_copy(slot, batch)
        cur.wait_event(self.loaded[slot])
        out = self.slots[slot].replay()
        if self.after_step is not None:
            self.after_step(self.slots[slot])
        self.result_host[slot].copy_(out.detach().sum().reshape(1), non_blocking=True)  # D2H of the step's result
        self.done[slot].record(cur)
        prev = None
        if self.pending is not None:
            pslot = self.pending
            self.done[pslot].synchronize()
            prev = float(self.result_host[pslot][0])
        self.pending = slot
        self.i += 1
        return prev

    def flush(self):
        if self.pending is None:
            return None
        self.done[self.pending].synchronize()
        v = float(self.result_host[self.pending][0])
        self.pending = None
        return v
