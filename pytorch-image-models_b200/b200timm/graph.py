"""CUDA-graph replay of a fixed-shape forward: the ~90 kernel launches of one ViT-B step (ctypes calls into
libb200timm + tensor-map encodes on the host) are captured once and replayed with a single cudaGraphLaunch, so the
step is never host-bound and inter-kernel gaps disappear.  Buffers are static: inputs are either copied into the
captured input tensor (`__call__`) or the graph is captured directly on a caller-owned resident buffer
(`capture_on`)."""
import torch


class GraphedForward:
    def __init__(self, fn, example_input, warmup=3):
        self.fn = fn
        self.static_in = example_input.clone()
        self.graph, self.static_out = self._capture(self.static_in, warmup)

    def _capture(self, x, warmup):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):   # warm-up outside capture: lazy cudaFuncSetAttribute calls, allocator growth
                self.fn(x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            out = self.fn(x)
        return g, out

    def capture_on(self, resident_input, warmup=1):
        """A second graph that reads `resident_input` in place (no staging copy); returns a zero-arg callable."""
        g, out = self._capture(resident_input, warmup)

        def replay():
            g.replay()
            return out
        replay.graph = g
        return replay

    def __call__(self, x):
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
