"""CUDA-graph capture of one micro-step (training_step + backward) so the ~340 (RN50) / ~950 (RN152) kernel
launches of a step are submitted with one cudaGraphLaunch instead of one Python -> ctypes call each.

The captured work is exactly what ``model.training_step(batch, i)`` followed by ``(loss * scale).backward()``
enqueues on the stream; the optimiser step stays outside (its learning rate / step count change per step)."""
import os
from typing import Dict

import torch


class GraphedStep:
    def __init__(self, model, example_batch: Dict[str, torch.Tensor], grad_scale: float = 1.0, warmup: int = 2,
                 share_inputs: bool = False):
        """share_inputs: use the example batch's own (device) tensors as the captured input buffers instead of
        clones -- a producer that writes the next batch straight into them (e.g. the GPU augmentation's `out=`)
        then needs no copy per step."""
        dev = model.engine.device
        self.model = model
        if share_inputs:
            assert all(v.device == dev for v in example_batch.values() if torch.is_tensor(v))
            self.static = {k: v for k, v in example_batch.items() if torch.is_tensor(v)}
        else:
            self.static = {k: v.to(dev).clone() for k, v in example_batch.items() if torch.is_tensor(v)}
        self.grad_scale = grad_scale
        self.graph = torch.cuda.CUDAGraph()
        model.train()
        # The warm-up below runs real steps; gradients are zeroed by the caller afterwards, but BatchNorm running
        # statistics / num_batches_tracked would keep the extra momentum updates (also right after a resume, when they
        # were just restored from a checkpoint): snapshot them here and put them back after the warm-up.
        eng = model.engine
        bn_state = [(b, b.clone()) for m in model.modules() for n, b in m._buffers.items()
                    if b is not None and n in ("running_mean", "running_var")]
        nbt_state = eng.nbt.clone() if eng.nbt is not None else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):  # allocator / lazy-init warm-up outside the capture
                out = model.forward_backward(self.static, grad_scale)
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            for b, saved in bn_state:
                b.copy_(saved)
            if nbt_state is not None:
                eng.nbt.copy_(nbt_state)
        torch.cuda.synchronize()
        # (gradients: callers zero them before the first replay -- Trainer and bench do; nothing runs during capture)
        from . import _lib

        before = _lib.LAUNCHES
        # capture on a high-priority stream: the critical path (fprop / dgrad / BatchNorm chain) then outranks the
        # weight-gradient kernels the engine issues on its default-priority side stream
        cap = torch.cuda.Stream(priority=-1) if os.environ.get("PECLR_GRAPH_PRIORITY", "1") != "0" else None
        # data parallel: a second capture of the same step WITH the per-stage gradient all-reduces (NCCL collectives are
        # capturable; thread-local capture mode keeps NCCL's watchdog thread out of it).  The plain graph serves the
        # micro-steps inside an accumulation window, the syncing one closes the window.  Both share one memory pool.
        self.sync_graph = None
        self.overlap_sync = (model._dp_world() > 1 and os.environ.get("PECLR_OVERLAP_ALLREDUCE", "1") != "0")
        model.enable_overlapped_sync(False)
        with torch.cuda.graph(self.graph, stream=cap):
            out = model.forward_backward(self.static, grad_scale)
        self.kernels_per_replay = _lib.LAUNCHES - before
        self.out = {k: v.detach() for k, v in out.items()}
        if self.overlap_sync:
            import torch.distributed as dist

            try:
                model.enable_overlapped_sync(True)
                side2 = torch.cuda.Stream()
                side2.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side2):  # NCCL warm-up on the communication stream, outside any capture
                    model.forward_backward(self.static, grad_scale)
                    model.sync_gradients()
                torch.cuda.current_stream().wait_stream(side2)
                torch.cuda.synchronize()
                dist.barrier()
                with torch.no_grad():
                    for b, saved in bn_state:
                        b.copy_(saved)
                    if nbt_state is not None:
                        eng.nbt.copy_(nbt_state)
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, stream=cap, pool=self.graph.pool(), capture_error_mode="thread_local"):
                    out2 = model.forward_backward(self.static, grad_scale)
                model._pending_sync = False  # (the join is inside the captured graph)
                self.sync_graph, self.sync_out = g2, {k: v.detach() for k, v in out2.items()}
            finally:
                model.enable_overlapped_sync(False)

    def matches(self, batch: Dict[str, torch.Tensor]) -> bool:
        return all(k in batch and batch[k].shape == v.shape for k, v in self.static.items())

    def __call__(self, batch: Dict[str, torch.Tensor], sync: bool = False) -> Dict[str, torch.Tensor]:
        """Copies the batch into the captured input buffers and replays the step.  The returned tensors are the
        graph's static outputs (valid until the next replay).  sync=True (the micro-step that closes an accumulation
        window, data parallel): replays the capture that also all-reduces the gradients stage by stage -- the caller
        must then NOT call model.sync_gradients() for this step (self.synced tells)."""
        from . import _lib

        for k, dst in self.static.items():
            if k not in batch:
                continue  # (extra keys of the example batch the model does not read)
            src = batch[k]
            if src.shape != dst.shape:
                raise ValueError(f"GraphedStep was captured for {k} of shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.synced = bool(sync and self.sync_graph is not None)
        if self.synced:
            self.sync_graph.replay()
            out = self.sync_out
        else:
            self.graph.replay()
            out = self.out
        _lib.LAUNCHES += self.kernels_per_replay
        self.model.train_metrics = dict(out)
        return out
