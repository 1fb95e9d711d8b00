"""Host -> device staging of mod_dict batches.

run_training_4m.py:712-716 moves every tensor of the loader's (pinned) batch with `.to(device, non_blocking=True)` on the
compute stream, so the copy of step i+1 waits behind the kernels of step i.  DevicePrefetcher issues the same copies on a
side stream from a background thread, `depth` batches ahead: the 80 MB mod7 batch (an fp32 RGB image per sample) rides under
the previous step's compute instead of in front of the next one.  The thread matters on hosts where the copy call itself
blocks (pageable or slowly-pinned memory: two pool boxes moved the batch at ~3 GB/s with the calling thread stuck in the
copy, starving the launch queue); the stream alone only helps when the call returns immediately.  Same tensors, same bytes,
no change to what the model sees."""
import queue
import threading

import torch

_END = object()


class DevicePrefetcher:
    """Iterate `batches` (an iterable of {modality: {key: CPU tensor, ideally pinned}}) as device-resident mod_dicts."""

    def __init__(self, batches, device, depth=2, stream=None):
        """stream: reuse a side stream across several prefetchers (the caching allocator keeps one pool per stream: a fresh stream pays
        cudaMalloc for every staged tensor again)."""
        self.batches = batches
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:        # the worker thread pins itself to an explicit device
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.depth = max(1, int(depth))
        self.stream = stream if stream is not None else torch.cuda.Stream(device=self.device)

    def _stage(self, host_batch):
        with torch.cuda.stream(self.stream):
            dev = {m: {k: v.to(self.device, non_blocking=True) for k, v in d.items()} for m, d in host_batch.items()}
            ev = self.stream.record_event()
        return dev, ev

    def _worker(self, q, stop):
        try:
            torch.cuda.set_device(self.device)
            for hb in self.batches:
                if stop.is_set():
                    return
                item = self._stage(hb)
                while not stop.is_set():
                    try:
                        q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
            item = _END
        except BaseException as e:          # surfaced in the consumer thread
            item = e
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                break
            except queue.Full:
                continue

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        t = threading.Thread(target=self._worker, args=(q, stop), daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is _END:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield self._hand_over(item)
        finally:
            stop.set()

    def _hand_over(self, staged):
        dev, ev = staged
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for d in dev.values():                      # the allocator must not recycle these blocks for the side stream while the
            for v in d.values():                    # compute stream still reads them
                v.record_stream(cur)
        return dev
