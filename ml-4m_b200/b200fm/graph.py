"""Whole-step CUDA graph for the 4M train step (forward + backward + gradient all-reduce + AdamW).

Why: one 4M-B step is ~830 kernel launches issued from Python (~20-30 ms of host time on the pool's hosts, ~100 ms for 4M-L's
48 blocks) against ~33 ms of GPU time -- the host, not the kernels, is what the step waits for.  The reference answers launch
overhead with `torch.compile`; here the step is captured ONCE with CUDA stream capture and replayed with one `cudaGraphLaunch`.

What makes the step replayable (everything that used to depend on the host is data now):
  * the Python-`random` shuffle of the decoder modalities (reference fm.py:306) is drawn on the host exactly like the reference does
    and uploaded as a device permutation (`b200fm_select_plan_ordered`);
  * the masked-token head works with device-side per-modality row counts (`FourM.static_head`, `b200fm_gemm_bf16_dyn`);
  * AdamW reads lr / bias corrections from device memory (`FusedAdamW(capturable=True)`), so schedules keep working;
  * the gradient all-reduce kernels take their sequence numbers from a device counter (b200fm.parallel.GradSync);
  * the batch is copied into static input buffers (host->device directly when the caller hands over pinned CPU tensors).

Usage (replaces the body of the reference's train loop, run_training_4m.py:712-745):

    step = GraphedTrainStep(model_or_GradSync, optimizer, num_encoder_tokens=128, num_decoder_tokens=128)
    for batch in loader:
        for g in optimizer.param_groups: g["lr"] = schedule(it)
        loss, mod_loss, grad_norm = step(batch)          # device tensors, valid until the next call
"""
import random

import torch

from . import functional as BF
from . import lib


class GraphedTrainStep:
    def __init__(self, net, optimizer, num_encoder_tokens, num_decoder_tokens, loss_type="mod", eager_steps=2, grad_norm=True):
        self.net = net
        self.model = net.module if hasattr(net, "module") else net
        self.opt = optimizer
        if not getattr(optimizer, "capturable", False):
            raise ValueError("GraphedTrainStep needs FusedAdamW(capturable=True): the step count / learning rate must live on the device")
        self.N, self.M, self.loss_type = num_encoder_tokens, num_decoder_tokens, loss_type
        self.eager_left = max(1, eager_steps)       # first calls run eagerly (allocator / caches / GradSync warm up), then capture
        self.want_norm = grad_norm
        self.graph = None
        self.static_in = None
        self.static_out = None
        self.order_dev = None
        self.dec_mods = None
        self.model.static_head = True
        self.replays = 0
        self.kernel_calls_per_step = 0

    # ------------------------------------------------------------------ pieces
    def _params(self):
        return [p for g in self.opt.param_groups for p in g["params"]]

    def _step_body(self, batch):
        loss, mod_loss = self.net(batch, num_encoder_tokens=self.N, num_decoder_tokens=self.M, loss_type=self.loss_type)
        loss.backward()
        BF.join_side_wgrad()            # (weight-gradient GEMMs that ran on the side stream, see functional._tn_gemm)
        # the logged gradient norm (native_scaler.py:56-65) of the averaged gradients: accumulated by the AdamW kernels while they read them
        self.opt.track_grad_norm = bool(self.want_norm)
        self.opt.step()
        gnorm = self.opt.grad_norm() if self.want_norm else None
        self.opt.zero_grad(set_to_none=True)
        return loss.detach(), {k: v.detach() for k, v in mod_loss.items()}, gnorm

    def _draw_order(self, batch):
        """The reference's per-step shuffle (fm.py:306): one random.sample over the decoder modalities present in the batch."""
        dec_mods = [m for m in batch if m in self.model.decoder_embeddings]
        order = random.sample(dec_mods, len(dec_mods))
        return dec_mods, [dec_mods.index(m) for m in order]

    def _copy_in(self, batch):
        for m, d in batch.items():
            sd = self.static_in[m]
            for k, v in d.items():
                sd[k].copy_(v, non_blocking=True)

    # ------------------------------------------------------------------ call
    def __call__(self, batch):
        dev = next(self.model.parameters()).device
        dec_mods, perm = self._draw_order(batch)
        self.opt.prepare_step()
        if self.graph is None and self.eager_left > 0:
            # eager step on the caller's batch with the same code path (device-side order, static head)
            self.eager_left -= 1
            self.model._decoder_order_dev = torch.tensor(perm, dtype=torch.int32, device=dev)
            try:
                dbatch = {m: {k: v.to(dev, non_blocking=True) for k, v in d.items()} for m, d in batch.items()}
                return self._step_body(dbatch)
            finally:
                self.model._decoder_order_dev = None
        if self.graph is None:
            self._capture(batch, dec_mods, dev)
        if dec_mods != self.dec_mods:
            raise ValueError(f"GraphedTrainStep was captured for the modalities {self.dec_mods}; this batch has {dec_mods}")
        self._copy_in(batch)
        self.order_dev.copy_(torch.tensor(perm, dtype=torch.int32), non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_out

    def _capture(self, batch, dec_mods, dev):
        self.dec_mods = dec_mods
        self.static_in = {m: {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in d.items()} for m, d in batch.items()}
        self.order_dev = torch.zeros(len(dec_mods), dtype=torch.int32, device=dev)
        self.model._decoder_order_dev = self.order_dev
        torch.cuda.synchronize(dev)
        BF.reset_zero_arena()
        self.opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        calls0 = lib.CALLS["n"]
        import os
        side = torch.cuda.Stream(device=dev) if os.environ.get("B200FM_SIDE_WGRAD", "0") == "1" else None
        try:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                if side is not None:
                    BF.enable_side_wgrad(side)
                out = self._step_body(self.static_in)
        finally:
            BF.join_side_wgrad(disable=True)
            BF.reset_zero_arena()
            lib.set_option("sm_reserve", 0)
            self.model._decoder_order_dev = None     # eager calls of the model (evaluation) draw their own shuffle again
        self.kernel_calls_per_step = lib.CALLS["n"] - calls0      # C-ABI kernel launches baked into one replay
        self.graph, self.static_out = g, out

    def release(self):
        """Drop the graph (and its private memory pool); later calls run eagerly again until re-captured."""
        self.graph = None
        self.static_in = self.static_out = None
        self.model._decoder_order_dev = None
        self.eager_left = 1


class GraphedCall:
    """CUDA-graph replay of an inference call with tensor inputs and tensor outputs (e.g. `VQ.tokenize` on a fixed batch shape: ~190
    launches whose host issue time exceeds their GPU time).  One graph per input signature (shapes + dtypes); the first `eager_calls`
    calls of a signature run eagerly (weight-limb / bf16-shadow / allocator caches fill up), the next one is captured.

        tok = GraphedCall(lambda x: vq.tokenize(x))
        tokens = tok(images)                       # clones of the graph's output buffers (clone=False: valid until the next call)

    `inputs(*example)` returns the graph's static input buffers, so a loader can copy pinned host memory straight into them and call
    `replay(*example)`.  The callee must be free of host synchronisation and must not depend on host-side state that changes between
    calls (weights updated in place are fine: the graph reads them where they live; re-create the wrapper after replacing parameters)."""

    def __init__(self, fn, eager_calls=2, clone=True):
        self.fn, self.eager_calls, self.clone = fn, max(1, eager_calls), clone
        self._seen, self._graphs = {}, {}

    @staticmethod
    def _key(args):
        return tuple((tuple(a.shape), a.dtype, a.device.index) for a in args)

    def _get(self, args):
        key = self._key(args)
        ent = self._graphs.get(key)
        if ent is None:
            n = self._seen.get(key, 0)
            if n < self.eager_calls:
                self._seen[key] = n + 1
                return None
            static_in = [torch.empty_like(a) for a in args]
            for s, a in zip(static_in, args):
                s.copy_(a)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = self.fn(*static_in)
            ent = self._graphs[key] = (static_in, g, out)
        return ent

    def inputs(self, *example):
        ent = self._get(example)
        return None if ent is None else ent[0]

    def replay(self, *example):
        static_in, g, out = self._graphs[self._key(example)]
        g.replay()
        return out

    def __call__(self, *args):
        ent = self._get(args)
        if ent is None:
            with torch.no_grad():
                return self.fn(*args)
        static_in, g, out = ent
        for s, a in zip(static_in, args):
            if s.data_ptr() != a.data_ptr():
                s.copy_(a, non_blocking=True)
        g.replay()
        if not self.clone:
            return out
        return out.clone() if torch.is_tensor(out) else type(out)(o.clone() if torch.is_tensor(o) else o for o in out)
