"""Data-parallel gradient synchronisation for the 4M train step: the B200 replacement for
`torch.nn.parallel.DistributedDataParallel(model, device_ids=[gpu])` of the reference (run_training_4m.py:512).

Same contract as DDP -- identical replicas, one process per GPU, after `loss.backward()` every parameter's `.grad` holds the
MEAN over ranks of the local fp32 gradients -- built differently:

  * ONE flat fp32 gradient arena per rank.  The weight-gradient GEMMs, the embedding scatter and the LayerNorm backward write
    straight into it (`b200fm.functional` asks `claim(param)` for the destination), so there is no bucket copy-in / copy-out.
  * The arena is cut into chunks in (approximately) the order backward produces the gradients.  When the last parameter of a chunk
    has its gradient, the chunk is reduced on a side stream by `b200fm_allreduce_f32` (csrc/comm.cu): a two-shot all-reduce over
    NVLink peer memory (CUDA IPC) on a fixed, small number of CTAs, while the persistent GEMM / attention grids leave exactly
    that many SMs free (runtime option "sm_reserve").  Chunks are launched strictly in index order, so every rank issues the
    same sequence of reductions.
  * Nothing waits at the end of backward: the optimizer asks `wait(params)` per parameter group, so AdamW of the early chunks
    runs while the last chunk (the embedding tables, whose gradients only exist at the very end of backward) is still in flight.

Transports: "p2p" (the kernel above; CUDA, 2..8 ranks of one node) and "collective" (`torch.distributed.all_reduce` on the chunk
views: NCCL as an A/B baseline, gloo for the CPU tests of this file's host logic).
"""
import ctypes
import os
from contextlib import contextmanager

import torch
import torch.distributed as dist

from . import lib

_ALIGN = 64            # slot alignment in fp32 elements (256 B)
SMALL_NUMEL = 16384    # parameters up to this size share one pre-zeroed region (norm weights, biases, mod_emb, mask token)


class _Slot:
    __slots__ = ("sync", "param", "name", "offset", "numel", "shape", "chunk", "small", "claimed", "count", "pair_rows", "expected")

    def view(self):
        s = self.sync
        return s.arena[self.offset:self.offset + self.numel].view(self.shape)


class _CudaBlob:
    """Raw device memory exposed through __cuda_array_interface__ so torch can alias it without owning it."""

    def __init__(self, ptr, n_f32):
        self.__cuda_array_interface__ = dict(shape=(n_f32,), typestr="<f4", data=(ptr, False), version=3)


class _P2PTransport:
    """Arena in cudaMalloc memory, exported to the peers with CUDA IPC; reductions by csrc/comm.cu."""
    name = "p2p"

    def __init__(self, n_elems, device, group, n_ctas):
        self.group, self.n_ctas = group, n_ctas
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = lib.load()
        self.flag_elems = L.b200fm_comm_flag_bytes() // 4
        total = self.flag_elems + n_elems
        ptr = ctypes.c_void_p()
        lib.check(L.b200fm_comm_alloc(ctypes.c_longlong(total * 4), ctypes.byref(ptr)), "b200fm_comm_alloc")
        self.base = ptr.value
        self._blob = _CudaBlob(self.base, total)
        whole = torch.as_tensor(self._blob, device=device)
        self.arena = whole[self.flag_elems:]
        handle = (ctypes.c_ubyte * 64)()
        lib.check(L.b200fm_comm_ipc_export(ctypes.c_void_p(self.base), handle), "b200fm_comm_ipc_export")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        self.peer_base = []
        for r in range(self.world):
            if r == self.rank:
                self.peer_base.append(self.base)
                continue
            hb = (ctypes.c_ubyte * 64)(*allh[r].cpu().tolist())
            p = ctypes.c_void_p()
            lib.check(L.b200fm_comm_ipc_open(hb, ctypes.byref(p)), "b200fm_comm_ipc_open")
            self.peer_base.append(p.value)
        self._data = (ctypes.c_void_p * self.world)(*[b + self.flag_elems * 4 for b in self.peer_base])
        self._flags = (ctypes.c_void_p * self.world)(*self.peer_base)
        dist.barrier(group=group)

    def all_reduce(self, offset, n, seq, stream, seq_base=None, n_ctas=None):
        """seq (+ *seq_base, a device int32 the caller advances once per step) must grow with every call, identically on all ranks."""
        lib.call("b200fm_allreduce_f32_seq", self._data, self._flags, self.rank, self.world, offset, n, 1.0 / self.world, seq & 0xFFFFFFFF,
                 0 if seq_base is None else seq_base.data_ptr(), n_ctas or self.n_ctas, stream.cuda_stream)

    def close(self):
        L = lib.load()
        for r, b in enumerate(self.peer_base):
            if r != self.rank:
                L.b200fm_comm_ipc_close(ctypes.c_void_p(b))
        L.b200fm_comm_free(ctypes.c_void_p(self.base))
        self.peer_base = []


class _CollectiveTransport:
    """Arena in ordinary torch memory; reductions by torch.distributed (NCCL / gloo)."""
    name = "collective"

    def __init__(self, n_elems, device, group, n_ctas):
        self.group = group
        self.world = dist.get_world_size(group)
        self.arena = torch.zeros(n_elems, dtype=torch.float32, device=device)

    def all_reduce(self, offset, n, seq, stream, seq_base=None, n_ctas=None):
        v = self.arena[offset:offset + n]
        if v.is_cuda:
            with torch.cuda.stream(stream):
                dist.all_reduce(v, group=self.group)
                v.mul_(1.0 / self.world)
        else:
            dist.all_reduce(v, group=self.group)
            v.mul_(1.0 / self.world)

    def close(self):
        pass


class GradSync(torch.nn.Module):
    """Wrap `model` like DDP: `net = GradSync(model); loss = net(...); loss.backward(); opt.step()`."""

    def __init__(self, model, process_group=None, transport="auto", chunk_mb=48, n_ctas=None, wait_at_end=True,
                 broadcast_params=True, small_numel=SMALL_NUMEL, n_ctas_tail=None):
        super().__init__()
        self.module = model
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.wait_at_end = wait_at_end
        params = self._ordered_params(model)
        if not params:
            raise ValueError("GradSync: the model has no parameters that require grad")
        self.device = params[0][1].device
        cuda = self.device.type == "cuda"
        if transport == "auto":
            transport = os.environ.get("B200FM_COMM", "p2p" if cuda and 2 <= self.world <= 8 else "collective")
        # Measured on 2 B200s (4M-B, profiles/r2_bench_2gpu_*.json): the reductions have 2/3 of a step to hide in, yet INSIDE the step the
        # kernel runs several times slower than alone (the GEMMs saturate the L2 -> SM path), so the exposed time falls with the number of
        # CTAs: wide 6 CTAs on reserved SMs 3.4 ms, wide 12: 1.5 ms, slim 48: 2.4 ms, slim 96: 1.0 ms, slim 128: 0.9 ms.
        # "slim" all-reduce CTAs (option comm_slim / B200FM_COMM_SLIM=1): 128 threads, <= 64 registers -- co-resident with the persistent
        # backward kernels instead of on SMs reserved for them; the bytes in flight come from the CTA count
        self.slim = cuda and transport == "p2p" and lib.get_option("comm_slim") != 0
        if n_ctas is None:
            n_ctas = int(os.environ.get("B200FM_COMM_CTAS", "128" if self.slim else "12"))
        if n_ctas_tail is None:
            n_ctas_tail = int(os.environ.get("B200FM_COMM_CTAS_TAIL", "128" if self.slim else "64"))
        self.n_ctas, self.n_ctas_tail = n_ctas, max(n_ctas, n_ctas_tail)
        self.small_numel = small_numel
        self._layout(params, int(chunk_mb * (1 << 20) // 4))
        self.transport = (_P2PTransport if transport == "p2p" else _CollectiveTransport)(self.total, self.device, process_group, n_ctas)
        self.arena = self.transport.arena
        self.reserve_sms = n_ctas if (cuda and transport == "p2p" and not self.slim) else 0
        self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1) if cuda else None
        self._step = 0
        # step-dependent part of the all-reduce sequence numbers, kept ON THE DEVICE and advanced by a stream-ordered add at the start
        # of every step: a CUDA graph that captured the step replays with fresh sequence numbers without host involvement
        self._seq_base = torch.zeros(1, dtype=torch.int32, device=self.device) if cuda else None
        self._sync_enabled = True
        self._active = False
        self._callback_queued = False
        self._next_chunk = 0
        self._learned = False
        self.stats = dict(direct=0, copied=0, launches=0)
        for s in self.slots:
            s.param._b200fm_slot = s
            s.param.register_post_accumulate_grad_hook(self._make_hook(s))
        if broadcast_params and self.world > 1:
            with torch.no_grad():
                for _, p in params:
                    dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
                for b in model.buffers():
                    if b.is_floating_point():
                        dist.broadcast(b.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)

    # ------------------------------------------------------------------ layout
    @staticmethod
    def _ordered_params(model):
        """Unique trainable parameters in (approximate) gradient-ready order = reverse registration order (the proxy DDP uses),
        with fc1 / fc3 of every SwiGLU block adjacent (their weight gradient is ONE [2H, D] GEMM output)."""
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        order = list(reversed(named))
        partner = {}
        for mname, m in model.named_modules():
            f1, f3 = getattr(m, "fc1", None), getattr(m, "fc3", None)
            if isinstance(f1, torch.nn.Linear) and isinstance(f3, torch.nn.Linear) and f1.weight.shape == f3.weight.shape:
                partner[id(f1.weight)] = (f1.weight, f3.weight)
                partner[id(f3.weight)] = (f1.weight, f3.weight)
        names = {id(p): n for n, p in named}
        out, seen = [], set()
        for n, p in order:
            if id(p) in seen:
                continue
            if id(p) in partner and all(q.requires_grad for q in partner[id(p)]):
                for q in partner[id(p)]:
                    if id(q) not in seen:
                        seen.add(id(q))
                        out.append((names[id(q)], q))
                continue
            seen.add(id(p))
            out.append((n, p))
        return out

    def _layout(self, params, chunk_elems):
        self.slots, big, small = [], [], []
        for n, p in params:
            s = _Slot()
            s.sync, s.param, s.name, s.numel, s.shape = self, p, n, p.numel(), tuple(p.shape)
            s.small = p.numel() <= self.small_numel
            s.claimed, s.count, s.pair_rows, s.expected = False, 0, 0, 1
            (small if s.small else big).append(s)
        off = 0
        self.chunks = []            # [start, end, [slots]]
        cur = [0, 0, []]
        for i, s in enumerate(big):
            rows = s.shape[0]
            pad_rows = (rows + 7) // 8 * 8 if len(s.shape) == 2 else rows      # SwiGLU pairs: rows padded like the bf16 operand
            s.offset = off
            s.pair_rows = pad_rows
            size = (pad_rows * (s.numel // max(rows, 1)) if len(s.shape) == 2 else s.numel)
            # a following partner must start exactly pad_rows * cols after this slot: no extra alignment inside a pair
            nxt = big[i + 1] if i + 1 < len(big) else None
            paired = nxt is not None and self._is_pair(s, nxt)
            off += size if paired else (size + _ALIGN - 1) // _ALIGN * _ALIGN
            cur[2].append(s)
            cur[1] = off
            if not paired and cur[1] - cur[0] >= chunk_elems:
                self.chunks.append(cur)
                cur = [off, off, []]
        self.small_start = off
        for s in small:
            s.offset = off
            off += (s.numel + 3) // 4 * 4
            cur[2].append(s)
        off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        self.small_end = off
        cur[1] = off
        if cur[2]:
            self.chunks.append(cur)
        self.total = off
        for ci, c in enumerate(self.chunks):
            for s in c[2]:
                s.chunk = ci
        self.slots = big + small
        self._pending = [len(c[2]) for c in self.chunks]
        self._events = [None] * len(self.chunks)

    @staticmethod
    def _is_pair(a, b):
        return (a.name.endswith("fc1.weight") and b.name.endswith("fc3.weight") and a.name[:-len("fc1.weight")] == b.name[:-len("fc3.weight")]
                and a.shape == b.shape)

    # ------------------------------------------------------------------ destination views for the backward kernels
    def claim(self, param, zeroed=False):
        """A fresh view of `param`'s arena slot for a kernel to write its gradient into, or None (already claimed this step,
        outside a step, or -- with zeroed=True -- a slot that is not pre-zeroed).  The FIRST producer of a step gets the slot;
        later producers of the same parameter (tied tables, shared mod_emb) return ordinary tensors that autograd adds in place."""
        s = getattr(param, "_b200fm_slot", None)
        if s is None or s.sync is not self or not self._active or s.claimed:
            return None
        if zeroed and not s.small:
            return None
        s.claimed = True
        return s.view()

    def claim_pair(self, w1, w3, rows_padded):
        """[2 * rows_padded, cols] view over the adjacent fc1 / fc3 slots (SwiGLU weight-gradient GEMM output), or None."""
        a, b = getattr(w1, "_b200fm_slot", None), getattr(w3, "_b200fm_slot", None)
        if a is None or b is None or a.sync is not self or not self._active or a.claimed or b.claimed:
            return None
        cols = a.shape[1]
        if a.pair_rows != rows_padded or b.offset != a.offset + rows_padded * cols:
            return None
        a.claimed = b.claimed = True
        return self.arena[a.offset:a.offset + 2 * rows_padded * cols].view(2 * rows_padded, cols)

    # ------------------------------------------------------------------ step protocol
    def forward(self, *args, **kwargs):
        if torch.is_grad_enabled():
            self.begin_step()
        return self.module(*args, **kwargs)

    def begin_step(self):
        for s in self.slots:
            s.claimed, s.count = False, 0
        self._pending = [len(c[2]) for c in self.chunks]
        self._next_chunk = 0
        self._events = [None] * len(self.chunks)
        self._callback_queued = False
        self._active = True
        if self.small_end > self.small_start:
            self.arena[self.small_start:self.small_end].zero_()          # the kernels ACCUMULATE into these (dgamma, dbeta, mod_emb, ...)
        if self._seq_base is not None and self._sync_enabled:
            self._seq_base.add_(len(self.chunks))

    def _make_hook(self, slot):
        def hook(p):
            if not self._active:
                return
            if not self._callback_queued:
                self._callback_queued = True
                torch.autograd.Variable._execution_engine.queue_callback(self.finish)
                if self.reserve_sms and self._sync_enabled and self._learned:
                    lib.set_option("sm_reserve", self.reserve_sms)      # backward GEMM / attention grids leave room for the reductions
            slot.claimed = True          # whoever produced this gradient, a later producer must ADD to it, not overwrite the slot
            g = p.grad
            if g is None:
                return
            if g.data_ptr() != self.arena.data_ptr() + slot.offset * 4 or not g.is_contiguous():
                v = slot.view()
                v.copy_(g)
                p.grad = v
                self.stats["copied"] += 1
            else:
                self.stats["direct"] += 1
            slot.count += 1
            if slot.count == 1:
                self._pending[slot.chunk] -= 1
                if self._learned and self._sync_enabled:
                    self._launch_ready()
            elif self._learned and self._events[slot.chunk] is not None:
                raise RuntimeError(f"GradSync: {slot.name} received a second gradient after its chunk was reduced; parameters with several "
                                   "gradient producers must have them in every step (re-create GradSync after changing the graph)")
        return hook

    def _chunk_early_ok(self, ci):
        return all(s.expected <= 1 for s in self.chunks[ci][2])

    def _launch_ready(self):
        while self._next_chunk < len(self.chunks) and self._pending[self._next_chunk] == 0 and self._chunk_early_ok(self._next_chunk):
            self._launch(self._next_chunk)
            self._next_chunk += 1

    def _launch(self, ci, tail=False):
        start, end, _ = self.chunks[ci]
        seq = ci + 1 if self._seq_base is not None else self._step * len(self.chunks) + ci + 1
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()                                      # everything the compute stream has produced so far
            self.comm_stream.wait_event(ev)
            from . import functional as BF
            side = BF.side_stream()                          # weight gradients issued on the captured side stream (functional._tn_gemm)
            if side is not None:
                self.comm_stream.wait_stream(side)
            self.transport.all_reduce(start, end - start, seq, self.comm_stream, self._seq_base, self.n_ctas_tail if tail else self.n_ctas)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
            self._events[ci] = done
        else:
            self.transport.all_reduce(start, end - start, seq, None)
            self._events[ci] = True
        self.stats["launches"] += 1

    def finish(self):
        """End of backward (queued on the autograd engine by the first gradient hook): zero the slots of parameters that got no
        gradient, reduce every chunk that is still outstanding, give the reserved SMs back."""
        if not self._active:
            return
        for s in self.slots:
            if s.count == 0:
                if not s.small and not s.claimed:
                    s.view().zero_()
                if s.param.grad is None and self._sync_enabled:
                    s.param.grad = s.view()                  # DDP semantics: an unused parameter ends up with the mean of zeros
        if not self._learned:
            # (the autograd engine sums the gradients of a multi-producer parameter -- tied token tables, shared mod_emb -- in its
            # input buffer and runs the accumulation hook ONCE, when the sum is final; `expected` stays 1 unless a caller
            # accumulates a parameter several times per step by hand)
            for s in self.slots:
                s.expected = max(1, s.count)
        if self._sync_enabled:
            while self._next_chunk < len(self.chunks):
                self._launch(self._next_chunk, tail=self._learned)
                self._next_chunk += 1
            if self.reserve_sms:
                lib.set_option("sm_reserve", 0)
            if self.wait_at_end:
                self.wait()
        self._learned = True
        self._active = False
        self._step += 1

    def wait(self, params=None):
        """Make the current stream wait for the reductions that cover `params` (default: all)."""
        if self.comm_stream is None:
            return
        if params is None:
            chunks = range(len(self.chunks))
        else:
            chunks = sorted({p._b200fm_slot.chunk for p in params if getattr(p, "_b200fm_slot", None) is not None and p._b200fm_slot.sync is self})
        cur = torch.cuda.current_stream(self.device)
        for ci in chunks:
            ev = self._events[ci]
            if ev is not None and ev is not True:
                cur.wait_event(ev)

    def split_param_groups(self, groups, parts=4):
        """Split every optimizer param group into up to `parts` sub-groups along the chunk order (ranges of chunks holding about the
        same number of elements), early chunks first.  With `FusedAdamW.pre_group_hook = sync.wait` each sub-group then waits only
        for its own reductions: AdamW of the early ranges runs while the last chunks (the embedding tables, produced at the very
        end of backward) are still in flight."""
        sizes = [c[1] - c[0] for c in self.chunks]
        total, acc, part_of = sum(sizes), 0, []
        for sz in sizes:
            part_of.append(min(parts - 1, int(parts * acc / max(total, 1))))
            acc += sz
        out = []
        for g in groups:
            buckets = [[] for _ in range(parts + 1)]
            for p in g["params"]:
                s = getattr(p, "_b200fm_slot", None)
                buckets[part_of[s.chunk] if s is not None and s.sync is self else parts].append(p)
            out += [dict(g, params=b) for b in buckets if b]
        return out

    @contextmanager
    def no_sync(self):
        """Gradients stay local (no reduction), like DDP.no_sync(): for gradient accumulation and for timing the step without
        communication."""
        old = self._sync_enabled
        self._sync_enabled = False
        try:
            yield
        finally:
            self._sync_enabled = old

    def params_equal_across_ranks(self):
        """True when every rank holds bit-identical parameters (checksum all-reduce: max == min of a float64 sum and of an xor-free
        integer digest)."""
        with torch.no_grad():
            acc = torch.zeros(2, dtype=torch.float64, device=self.device)
            for s in self.slots:
                p = s.param.detach()
                acc[0] += p.double().sum()
                acc[1] += p.double().abs().sum()
            lo, hi = acc.clone(), acc.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            return bool(torch.equal(lo, hi))

    def close(self):
        for s in self.slots:
            if getattr(s.param, "_b200fm_slot", None) is s:
                del s.param._b200fm_slot
        self.transport.close()


def active_sync(param):
    """The GradSync that owns `param`'s gradient slot while a step is in flight, else None (hot-path helper for functional.py)."""
    s = getattr(param, "_b200fm_slot", None)
    if s is None:
        return None
    return s.sync if s.sync._active else None
