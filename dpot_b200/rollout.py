"""Pipelined autoregressive rollout (the loop of evaluate.py:192-208 / train_temporal.py:262-272
under no_grad): im = model(xx); pred[..., t] = im; xx = cat(xx[..., T_bundle:, :], im).

Everything is enqueued on one CUDA stream without host synchronisation.  The window is a RING in
time: logical frame t lives in slot (t + t0) % T, the model reads it through that offset
(dpot_rollout_step) and each step only overwrites the oldest T_bundle slots with the new frames:
the output-tail kernel stores them straight into the ring slot and into the preallocated
prediction tensor -- no torch.cat, no copy kernel, no copy of the surviving T - T_bundle frames."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class RolloutEngine:
    """use_graph: after one eager rollout (which also warms the library's one-time kernel attributes) the whole
    n_steps rollout -- ~50 kernel launches per step -- is captured into ONE CUDA graph and replayed; the buffers
    (window, prediction, workspace, packed weights) are fixed, so the recorded pointers stay valid.  The graph is
    dropped whenever the parameters change (the packed-weight arena is re-derived then)."""

    def __init__(self, model, batch: int, n_steps: int, device=None, use_graph: bool = False, want_cls: bool = False):
        self.model = model
        self.n_steps = n_steps
        self.use_graph = use_graph
        self._graph = None
        self._graph_key = None
        self._warm = False
        self.launches_per_run = 0      # library kernels per rollout (counted on the eager run; a graph replay launches the same)
        dev = device or next(model.parameters()).device
        R, T, Cc = model.img_size, model.in_timesteps, model.in_channels
        Tb, Co = model.out_timesteps, model.out_channels
        if Co != Cc:
            raise ValueError("autoregressive rollout needs out_channels == in_channels")
        self.win = torch.empty((batch, R, R, T, Cc), device=dev)
        self.im = torch.empty((batch, R, R, Tb, Co), device=dev)
        self.pred = torch.empty((batch, R, R, n_steps * Tb, Co), device=dev)
        # want_cls: also evaluate the classification head on every step, as the reference's model(xx) does
        # (evaluate.py:198 discards it); cls[s] = cls_pred of step s
        self.cls = torch.empty((n_steps, batch, model.n_cls), device=dev) if want_cls else None
        # the workspace is owned here: a captured graph holds its address, so it must not be shared with (or evicted
        # by) other batch sizes running through the model's engine
        with torch.cuda.device(dev):
            self.ws = model.engine().new_workspace(batch, dev)

    @torch.no_grad()
    def load(self, xx: torch.Tensor, non_blocking: bool = True) -> None:
        """Fill the window from xx[B,X,Y,T,C] (host-pinned or device) on the current stream."""
        self.win.copy_(xx, non_blocking=non_blocking)

    @torch.no_grad()
    def run(self, xx: Optional[torch.Tensor] = None, non_blocking: bool = True) -> torch.Tensor:
        """xx[B,X,Y,T,C] (host-pinned or device; None: the window was filled by load()) ->
        pred[B,X,Y,n_steps*T_bundle,C] on the device."""
        eng = self.model.engine()
        if xx is not None:
            self.load(xx, non_blocking)
        with torch.cuda.device(self.win.device):
            eng.refresh(self.win.device)          # (re-)pack the weights outside the counted / captured region
        if not self.use_graph:
            l0 = eng.lib.dpot_launch_count()
            self._steps(eng)
            self.launches_per_run = int(eng.lib.dpot_launch_count() - l0)
            return self.pred
        key = (eng._param_key(), eng.packed.data_ptr())
        if self._graph is not None and key != self._graph_key:
            self._graph, self._warm = None, False
        if not self._warm:                      # first rollout: eager (also the warm-up the capture needs)
            l0 = eng.lib.dpot_launch_count()
            self._steps(eng)
            self.launches_per_run = int(eng.lib.dpot_launch_count() - l0)
            self._warm, self._graph_key = True, (eng._param_key(), eng.packed.data_ptr())
            return self.pred
        if self._graph is None:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=self.win.device)
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                # thread_local: other threads (e.g. the NCCL watchdog of a multi-GPU job) may keep calling the CUDA API
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                    self._steps(eng)
            cur.wait_stream(side)
            self._graph = g
        self._graph.replay()
        return self.pred

    def _steps(self, eng) -> None:
        T, Tb = self.model.in_timesteps, self.model.out_timesteps
        t0 = 0
        for s in range(self.n_steps):
            # forward + window advance in one library call: the output tail writes the new frames straight into the
            # ring slot and the prediction tensor (dpot_rollout_step)
            eng.rollout_step(self.win, self.im, self.pred, t0, s, ws=self.ws,
                             want_cls=self.cls[s] if self.cls is not None else None)
            t0 = (t0 + Tb) % T


def rollout(model, xx: torch.Tensor, n_steps: int, engine: Optional[RolloutEngine] = None) -> torch.Tensor:
    eng = engine or RolloutEngine(model, xx.shape[0], n_steps, device=xx.device if xx.is_cuda else None)
    return eng.run(xx)
