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
