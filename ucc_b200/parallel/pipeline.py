"""Pipeline parallelism over send / recv (PP = p2p through two-member active-set broadcasts, the shape of
reference test/gtest/active_set/test_active_set.cc:165-184).  A stage owns a slice of the model; activations travel to the next
stage and gradients back with `ops.send / ops.recv` - tl/nvl's eager ring / rendezvous kernels for CUDA tensors, tl/shm for host
tensors.  Two schedules: GPipe (all forwards, then all backwards) and 1F1B (warm-up forwards, then one forward + one backward per
step, which bounds the activations a stage keeps alive by the pipeline depth instead of the number of micro-batches)."""
from __future__ import annotations

import torch

from .. import ops


class PipelineStage:
    """`module`: this rank's slice of the model.  `act_shape` / `dtype`: shape of ONE micro-batch of the activation this stage
    RECEIVES from the previous one (ignored on the first stage) - every stage passes what the next one declared.  Stage order =
    rank order of `comm`."""

    def __init__(self, module, act_shape=None, dtype=torch.float32, comm=None, device=None):
        self.module, self.comm = module, comm or ops.default_comm()
        self.rank, self.n = self.comm.rank, self.comm.size
        self.first, self.last = self.rank == 0, self.rank == self.n - 1
        self.act_shape, self.dtype = tuple(act_shape) if act_shape is not None else None, dtype
        self.device = device if device is not None else next(module.parameters()).device

    # activations use even tags, gradients odd ones (the host transport matches by tag; tl/nvl matches in post order per pair)
    def _forward(self, i, inputs, targets, loss_fn, n_micro, st):
        if self.first:
            x = inputs[i]
        else:
            x = torch.empty(self.act_shape, dtype=self.dtype, device=self.device)
            ops.recv(x, self.rank - 1, tag=2 * i, comm=self.comm)
            x.requires_grad_(True)
        y = self.module(x)
        if self.last:
            y = loss_fn(y, targets[i]) / n_micro
            st["loss"] = st["loss"] + y.detach()
        else:
            out = y.detach().contiguous()
            st["sends"].append((ops.send(out, self.rank + 1, tag=2 * i, comm=self.comm, async_op=True), out))
        st["live"][i] = (x, y)

    def _backward(self, i, st):
        x, y = st["live"].pop(i)
        if self.last:
            y.backward()
        else:
            g = torch.empty_like(y)
            ops.recv(g, self.rank + 1, tag=2 * i + 1, comm=self.comm)
            y.backward(g)
        if not self.first:
            gx = x.grad.contiguous()
            st["sends"].append((ops.send(gx, self.rank - 1, tag=2 * i + 1, comm=self.comm, async_op=True), gx))

    def run(self, n_micro, inputs=None, targets=None, loss_fn=None, schedule="1f1b"):
        """One training step over `n_micro` micro-batches.  First stage: `inputs[i]`; last stage: `targets[i]` and `loss_fn`.
        Parameter gradients accumulate (sum over micro-batches of grad(loss_i / n_micro)): the result equals the gradient of the
        mean loss of the unsplit model.  Returns the mean loss on the last stage, None elsewhere."""
        st = {"live": {}, "sends": [], "loss": torch.zeros((), device=self.device)}
        if schedule == "gpipe":
            for i in range(n_micro):
                self._forward(i, inputs, targets, loss_fn, n_micro, st)
            for i in range(n_micro):
                self._backward(i, st)
        elif schedule == "1f1b":
            warm = min(self.n - 1 - self.rank, n_micro)
            for i in range(warm):
                self._forward(i, inputs, targets, loss_fn, n_micro, st)
            for k in range(n_micro - warm):
                self._forward(warm + k, inputs, targets, loss_fn, n_micro, st)
                self._backward(k, st)
            for k in range(n_micro - warm, n_micro):
                self._backward(k, st)
        else:
            raise ValueError(schedule)
        for w, _keep in st["sends"]:      # sends are asynchronous: their buffers stay referenced until they are through
            w.wait()
        return st["loss"] if self.last else None
