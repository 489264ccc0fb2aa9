"""CUDA-graph replay of an SV model forward (inference): one capture per input shape, then every
batch is a device-side replay -- no per-kernel launch cost, and the sub-batches that
`fused.chunked` puts on separate streams really overlap (eager enqueueing serialises them on the host).

    net = svnet_b200.SV_DGCNN_CLS(args, 40).cuda().eval(); net.load_state_dict(...)
    fast = svnet_b200.GraphedForward(net, example_batch)      # same signature as net(...)
    logits = fast(batch)                                       # valid until the next call

Shapes are static: a batch of another shape raises (build a second GraphedForward for it).

``epilogue`` (optional callable ``y -> anything``) runs inside the capture right after the forward, on the
capture stream: multi-GPU callers pass the NCCL all-gather of the outputs here
(``lambda y: dist.all_gather_into_tensor(all_y, y)``) so that the collective is a node of the same graph
instead of a host-launched kernel behind every replay.  It is also run in the warm-up passes, so the
communicator exists before the capture starts.
"""
import torch


class GraphedForward:
    def __init__(self, model, *example_inputs, warmup=2, epilogue=None):
        if model.training:
            raise RuntimeError("GraphedForward captures the inference path: call model.eval() first")
        for t in example_inputs:
            if not t.is_cuda:
                raise RuntimeError("GraphedForward needs CUDA example inputs (there is no CPU path)")
        self.model = model
        self.static_in = [t.detach().clone() for t in example_inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):          # allocator / lazy weight packing settle before the capture
                y = model(*self.static_in)
                if epilogue is not None:
                    epilogue(y)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _native as nv
        self.graph = torch.cuda.CUDAGraph()
        l0 = nv.LAUNCHES[0]
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = model(*self.static_in)
            if epilogue is not None:
                epilogue(self.static_out)
        self.kernels_per_replay = nv.LAUNCHES[0] - l0     # kernels of this library inside one replay

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_in):
            raise TypeError("expected %d inputs, got %d" % (len(self.static_in), len(inputs)))
        for dst, src in zip(self.static_in, inputs):
            if tuple(src.shape) != tuple(dst.shape) or src.dtype != dst.dtype:
                raise ValueError("GraphedForward was captured for %s %s, got %s %s"
                                 % (tuple(dst.shape), dst.dtype, tuple(src.shape), src.dtype))
            dst.copy_(src, non_blocking=True)        # device-to-device, or pinned host-to-device
        self.graph.replay()
        return self.static_out
