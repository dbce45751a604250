"""CUDA-graph capture of a launch-bound fitting step.

Small scenes (BASELINE configs C1..C4: a few thousand Gaussians, 128^2..512^2 pixels) spend 0.3-0.9 ms in the path's
kernels per step but ~1-2 ms on the host: ~40 launches through Python + ctypes, each a few microseconds of GPU work.
`GraphedStep` captures the whole step -- renderer forward (pack, binning, trace, select, blend), gather-blend, loss and
the fused backward -- into ONE CUDA graph and replays it with a single launch.

What makes the step capturable is the speculative binning of `_C.BinPlan`: inside a capture no host code may wait for
the device, so the scratch sizes that the exact pass reads back after voge_bin_count (tile-list entries, hit slots)
come from the previous eager call of the same shape, the kernels never step outside those capacities
(include/voge_b200.h "Speculative scratch"), and the capacity check of `_C.bins_valid` is evaluated by the device on
every replay into a sticky violation flag.  `GraphedStep.__call__` replays, reads the flag and -- if a scene outgrew
its capacities (by more than the plan's 12.5 % slack) -- runs the step eagerly (which refreshes the plans through the
exact pass) and captures it again, so results are always those of the eager step.

Requirements on `step_fn` (the usual ones of CUDA-graph capture): it is re-runnable (zeroes the gradients it
accumulates into, e.g. `GradientBucket.zero()`), its inputs and parameters keep their storage between calls
(optimizers that update in place are fine and belong OUTSIDE the graph, like the gradient all-reduce), gradients
accumulate into existing `.grad` tensors, and it does not synchronise with the host.  The tensors it returns are
static: every replay overwrites them (they are returned detached).  References to losses / images of EARLIER eager
steps must be dropped before capturing: they keep the parameters' AccumulateGrad nodes bound to the stream of that
eager step (usually the legacy default stream), which a capture can not depend on.
"""
import torch

from . import _C, _lib


def _detach(x):
    if torch.is_tensor(x):
        return x.detach()
    if isinstance(x, (tuple, list)):
        return type(x)(_detach(v) for v in x)
    if isinstance(x, dict):
        return {k: _detach(v) for k, v in x.items()}
    return x


class _CaptureState(object):
    def __init__(self, device):
        self.violation = torch.zeros((), dtype=torch.bool, device=device)


class GraphedStep(object):
    def __init__(self, step_fn, device=None, warmup=3, validate=True):
        """step_fn() -> tensor(s).  `warmup` eager calls on a side stream (>= 2: the first call of a shape takes the
        exact binning pass and leaves the BinPlan the capture needs).  validate=False skips the per-replay read of
        the violation flag (check `violated()` yourself, e.g. once per epoch)."""
        self.step_fn = step_fn
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.validate = bool(validate)
        self.warmup = max(int(warmup), 2)
        self.recaptures = 0
        self.replays = 0
        self._flag_host = torch.zeros((), dtype=torch.bool).pin_memory()
        self._capture()

    def _capture(self):
        dev = self.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.step_fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.state = _CaptureState(dev)
        self.graph = torch.cuda.CUDAGraph()
        timer, _lib.kernel_timer = _lib.kernel_timer, None
        l0 = _lib.launch_count
        _C.capture_state = self.state
        try:
            with torch.cuda.graph(self.graph):
                out = self.step_fn()
            # static outputs without their autograd graph: a loss that kept the captured graph's AccumulateGrad nodes
            # alive would tie later eager backward passes to the capture stream
            self.outputs = _detach(out)
            del out
        finally:
            _C.capture_state = None
            _lib.kernel_timer = timer
        self.launches_per_replay = _lib.launch_count - l0       # C-ABI kernel launches inside the graph

    def violated(self):
        """True if a replay since the last capture found a scene larger than its scratch (host sync)."""
        self._flag_host.copy_(self.state.violation)
        return bool(self._flag_host)

    def __call__(self):
        self.graph.replay()
        self.replays += 1
        _lib.launch_count += self.launches_per_replay
        if self.validate and self.violated():
            # the scene outgrew the capacities baked into the graph: the eager step takes the exact pass (correct
            # results, fresh plans); then capture again with the new sizes
            self.recaptures += 1
            self._capture()
            self.graph.replay()
            _lib.launch_count += self.launches_per_replay
        return self.outputs
