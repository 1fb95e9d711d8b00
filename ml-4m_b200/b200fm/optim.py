"""Fused AdamW on the b200fm kernels.

Same update rule and defaults as torch.optim.AdamW as configured by the reference (fourm/utils/optim_factory.py:239-240:
betas (0.9, 0.95), weight decay 0.05 with no decay on norm / bias / 1-D tensors via param groups).  One multi-tensor launch
per param group: read p, g, m, v / write p, m, v (+ the bf16 weight shadow the GEMMs consume) in a single pass."""
import numpy as np
import torch

from . import functional as BF
from . import lib, ops


class FusedAdamW(torch.optim.Optimizer):
    """capturable=True keeps the per-step scalars (learning rate, bias corrections) in device memory: call `prepare_step()` once per
    step (it advances the step count and uploads 3 floats per group), then `step()` -- or replay a CUDA graph that captured `step()`
    (b200fm.graph.GraphedTrainStep).  All parameters then share one step count."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.capturable = capturable
        self._cap_step = 0
        self._hyper = {}        # group index -> device fp32 [3] = {lr, 1 - beta1^t, sqrt(1 - beta2^t)}
        self._chunk = None
        self._tables = {}       # group index -> dict(device table, chunk maps, pinned staging ring)
        self.pre_group_hook = None   # callable(params): e.g. GradSync.wait -- make the stream wait for this group's reduced gradients
        # track_grad_norm: the AdamW kernels also accumulate sum(g^2) of everything they update (stragglers on the single-tensor kernel
        # excluded); `grad_norm()` is the global L2 norm of the gradients of the last step() -- what run_training_4m.py logs -- for free
        self.track_grad_norm = False
        self._gnorm_sq = None

    def _group_tables(self, gi, tensors):
        """Device-side pointer table + CTA->chunk maps for one param group.  The chunk maps depend on the tensor sizes only
        (built once); the pointer table is re-uploaded when a pointer moved (set_to_none gradients are re-allocated every
        step) through a small ring of pinned staging buffers, each guarded by the event of its last upload."""
        if self._chunk is None:
            self._chunk = lib.load().b200fm_adamw_chunk_elems()
        tab = np.array(tensors, dtype=np.int64)                       # [n_t, 6]: p, g, m, v, shadow, n
        ent = self._tables.get(gi)
        sizes = tab[:, 5]
        if ent is None or ent["sizes"].shape != sizes.shape or not np.array_equal(ent["sizes"], sizes):
            dev = torch.device("cuda", torch.cuda.current_device())
            n_chunks = (sizes + self._chunk - 1) // self._chunk
            ct = np.repeat(np.arange(len(sizes), dtype=np.int32), n_chunks)
            starts = np.cumsum(n_chunks) - n_chunks
            co = (np.arange(int(n_chunks.sum()), dtype=np.int64) - np.repeat(starts, n_chunks)) * self._chunk
            ent = dict(sizes=sizes.copy(), ct=torch.from_numpy(ct).to(dev), co=torch.from_numpy(co).to(dev), n_chunks=int(n_chunks.sum()),
                       table=torch.empty(tab.shape, dtype=torch.int64, device=dev), last=None, ring=[], slot=0)
            for _ in range(4):
                ent["ring"].append([torch.empty(tab.shape, dtype=torch.int64).pin_memory(), None])
            ent["graph_host"] = torch.empty(tab.shape, dtype=torch.int64).pin_memory()     # source of a captured upload (never reused)
            self._tables[gi] = ent
        if torch.cuda.is_current_stream_capturing():
            # the upload becomes a memcpy node of the graph: it needs a pinned source of its own that nothing overwrites later
            host = ent["graph_host"]
            host.numpy()[...] = tab
            ent["table"].copy_(host, non_blocking=True)
            ent["last"] = None                                        # eager steps after the capture re-upload their own table
        elif ent["last"] is None or not np.array_equal(ent["last"], tab):
            host, ev = ent["ring"][ent["slot"]]
            if ev is not None:
                ev.synchronize()                                      # four uploads ago: long done
            host.numpy()[...] = tab
            ent["table"].copy_(host, non_blocking=True)
            ev = ent["ring"][ent["slot"]][1] = ev if ev is not None else torch.cuda.Event()
            ev.record()
            ent["slot"] = (ent["slot"] + 1) % 4
            ent["last"] = tab
        return ent

    def prepare_step(self):
        """capturable mode: advance the step count and refresh the device-side scalars of every group (tiny pageable H2D copies,
        stream-ordered before the kernels that read them)."""
        import math
        self._cap_step += 1
        t = self._cap_step
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            dev = self._hyper.get(gi)
            if dev is None:
                p0 = group["params"][0]
                dev = self._hyper[gi] = torch.zeros(3, device=p0.device, dtype=torch.float32)
            dev.copy_(torch.tensor([float(group["lr"]), 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t)], dtype=torch.float32), non_blocking=True)

    def state_dict(self):
        if self.capturable:      # replays do not run the Python loop that stamps the per-parameter step
            for st in self.state.values():
                if st:
                    st["step"] = self._cap_step
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        gn = None
        if self.track_grad_norm:
            dev = next((p.device for g in self.param_groups for p in g["params"] if p.grad is not None), None)
            if dev is not None:
                if self._gnorm_sq is None or self._gnorm_sq.device != dev:
                    self._gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
                self._gnorm_sq.zero_()
                gn = self._gnorm_sq
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            if self.pre_group_hook is not None:
                self.pre_group_hook(group["params"])
            tensors, extra_casts, step_no, updated, keepalive = [], [], None, [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32:
                    raise TypeError(f"FusedAdamW updates fp32 master parameters only (got {p.dtype}); the reference keeps fp32 "
                                    "parameters under autocast (run_training_4m.py:512), use torch.optim.AdamW for anything else")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = self._cap_step if self.capturable else st["step"] + 1
                updated.append(p)
                if step_no is None:
                    step_no = st["step"]
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if g.dtype != torch.float32:
                    g = g.float()
                if g is not p.grad:
                    keepalive.append(g)                  # local: must outlive the launch, must not end up in state_dict()
                if st["step"] != step_no:
                    # stragglers (a parameter that joined later has a different step count): single-tensor kernel
                    views = BF.shadow_views(p)
                    ops.adamw_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"],
                                   group["weight_decay"], st["step"], grad_scale, shadow=views[0] if views else None)
                    extra_casts += [(p, v) for v in views[1:]]
                    continue
                views = BF.shadow_views(p)
                tensors.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                views[0].data_ptr() if views else 0, p.numel()))
                extra_casts += [(p, v) for v in views[1:]]
            if tensors and self.capturable:
                if gi not in self._hyper:
                    raise RuntimeError("FusedAdamW(capturable=True): call prepare_step() before step()")
                ent = self._group_tables(gi, tensors)
                if gn is not None:
                    lib.call("b200fm_adamw_multi_gnorm", ent["table"].data_ptr(), ent["ct"].data_ptr(), ent["co"].data_ptr(), ent["n_chunks"], 0.0, float(b1),
                             float(b2), float(group["eps"]), float(group["weight_decay"]), 1, float(grad_scale), self._hyper[gi].data_ptr(), gn.data_ptr(),
                             ops._stream())
                else:
                    lib.call("b200fm_adamw_multi_dev", ent["table"].data_ptr(), ent["ct"].data_ptr(), ent["co"].data_ptr(), ent["n_chunks"], float(b1),
                             float(b2), float(group["eps"]), float(group["weight_decay"]), float(grad_scale), self._hyper[gi].data_ptr(), ops._stream())
            elif tensors:
                ent = self._group_tables(gi, tensors)
                if gn is not None:
                    lib.call("b200fm_adamw_multi_gnorm", ent["table"].data_ptr(), ent["ct"].data_ptr(), ent["co"].data_ptr(), ent["n_chunks"], float(group["lr"]),
                             float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), int(step_no), float(grad_scale), None, gn.data_ptr(),
                             ops._stream())
                else:
                    lib.call("b200fm_adamw_multi", ent["table"].data_ptr(), ent["ct"].data_ptr(), ent["co"].data_ptr(), ent["n_chunks"], float(group["lr"]),
                             float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), int(step_no), float(grad_scale),
                             ops._stream())
            for p, view in extra_casts:          # a weight mirrored in more than one operand buffer
                ops.cast_bf16(p.data, view)
            # the kernels wrote through raw pointers: bump the version counters (other version-keyed caches must notice) and
            # re-stamp the bf16 mirrors that were refreshed in the same pass
            BF.mark_updated(updated)
        return loss


def _fused_grad_norm(self):
    """Global L2 norm of the (scaled) gradients the last step() consumed, as a 0-d device tensor (track_grad_norm=True)."""
    if self._gnorm_sq is None:
        raise RuntimeError("FusedAdamW.grad_norm(): set track_grad_norm = True before step()")
    return self._gnorm_sq.sqrt().reshape(())


FusedAdamW.grad_norm = _fused_grad_norm


def param_groups_like_reference(model, weight_decay=0.05):
    """The decay / no-decay split of fourm/utils/optim_factory.py:111-168 (`get_parameter_groups` as `create_optimizer` calls it,
    :188-199): no weight decay for names containing "norm." / ".norm", ending in ".bias", ".lookup_table_weight" or ".gamma", and
    for the model's own `no_weight_decay()` list; everything else decays.  (No layer-wise lr scaling: 4M pre-training does not use it.)"""
    skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else set()
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if ("norm." in name or ".norm" in name or name.endswith(".bias") or name.endswith(".lookup_table_weight")
                or name.endswith(".gamma") or name in skip):
            no_decay.append(p)
        else:
            decay.append(p)
    return [dict(params=decay, weight_decay=weight_decay), dict(params=no_decay, weight_decay=0.0)]


class FusedModelEma(torch.nn.Module):
    """Drop-in for the reference's `ModelEmaV2` (fourm/utils/timm/model_ema.py:84-127; `run_training_vqvae.py:683` creates it, `:1171`
    calls `update(model)` every step): `.module` is an eval-mode copy of the model whose state_dict entries track
    `decay * ema + (1 - decay) * model`.  All fp32 entries (parameters and buffers) are updated by ONE multi-tensor kernel launch
    (b200fm_ema_multi; the reference issues three element-wise kernels per tensor, ~600 launches for a ViT-B/ViT-B VQ-VAE), bit-identical
    to the reference's expression; the few non-fp32 entries (integer buffers) go through the reference's torch expression.
    device: only None / the model's own device ('' like the reference's default); an EMA on another device would defeat the kernel."""

    def __init__(self, model, decay=0.9999, device=None, resume=''):
        super().__init__()
        import copy
        if device:
            raise NotImplementedError("FusedModelEma keeps the average on the model's device; use the reference's ModelEmaV2 for device='cpu'")
        if resume:
            raise NotImplementedError("FusedModelEma: load the checkpoint's 'state_dict_ema' into `.module` yourself (module.load_state_dict)")
        self.module = copy.deepcopy(model)
        self.module.eval()
        self.decay = decay
        self.device = device
        self.ema_has_module = hasattr(self.module, 'module')
        self._plan = None

    def _build_plan(self, model):
        ema_vals, model_vals = list(self.module.state_dict().values()), list(model.state_dict().values())
        if len(ema_vals) != len(model_vals):
            raise ValueError("FusedModelEma.update: the model's state_dict does not match the averaged copy")
        fused, rest = [], []
        for e, m in zip(ema_vals, model_vals):
            if e.shape != m.shape:
                raise ValueError("FusedModelEma.update: state_dict entries differ in shape")
            ok = e.is_cuda and m.is_cuda and e.dtype == torch.float32 and m.dtype == torch.float32 and e.is_contiguous() and m.is_contiguous() and e.numel() > 0
            (fused if ok else rest).append((e, m))
        chunk = lib.load().b200fm_adamw_chunk_elems()
        plan = dict(fused=fused, rest=rest, n_chunks=0)
        if fused:
            dev = fused[0][0].device
            sizes = np.array([e.numel() for e, _ in fused], dtype=np.int64)
            n_chunks = (sizes + chunk - 1) // chunk
            ct = np.repeat(np.arange(len(sizes), dtype=np.int32), n_chunks)
            starts = np.cumsum(n_chunks) - n_chunks
            co = (np.arange(int(n_chunks.sum()), dtype=np.int64) - np.repeat(starts, n_chunks)) * chunk
            plan.update(ct=torch.from_numpy(ct).to(dev), co=torch.from_numpy(co).to(dev), n_chunks=int(n_chunks.sum()), table=None, ptrs=None)
        # parameters of the averaged copy whose bf16 operand mirrors must be re-derived after the raw-pointer update
        plan["ema_params"] = [p for p in self.module.parameters()]
        return plan

    def _table(self, plan):
        ptrs = [(e.data_ptr(), m.data_ptr()) for e, m in plan["fused"]]
        if plan["table"] is None or plan["ptrs"] != ptrs:            # state_dict tensors are stable across steps: built once
            tab = np.zeros((len(ptrs), 6), dtype=np.int64)            # b200fm_adamw_tensor rows: p, g, m, v, shadow, n
            for i, ((e, m), (pe, pm)) in enumerate(zip(plan["fused"], ptrs)):
                tab[i] = (pe, pm, 0, 0, 0, e.numel())
            plan["table"] = torch.from_numpy(tab).to(plan["ct"].device)
            plan["ptrs"] = ptrs
        return plan["table"]

    @torch.no_grad()
    def update(self, model):
        if self._plan is None:
            self._plan = self._build_plan(model)
        plan = self._plan
        if plan["n_chunks"]:
            table = self._table(plan)
            lib.call("b200fm_ema_multi", table.data_ptr(), plan["ct"].data_ptr(), plan["co"].data_ptr(), plan["n_chunks"], float(self.decay), float(1. - self.decay), ops._stream())
        for e, m in plan["rest"]:
            e.copy_(self.decay * e + (1. - self.decay) * m)
        BF.mark_updated(plan["ema_params"])                           # version-keyed caches (bf16 mirrors) of the copy are stale now

    @torch.no_grad()
    def set(self, model):
        for e, m in zip(self.module.state_dict().values(), model.state_dict().values()):
            e.copy_(m)
        BF.mark_updated(list(self.module.parameters()))
