"""Optimiser / scheduler factories and the PNG saver (drop-in for the reference's util/util.py:8-59)."""
from __future__ import annotations

from pathlib import Path

import torch
from torch.optim import lr_scheduler


def get_scheduler(optimizer, lr_policy, n_epochs=None, n_epochs_decay=None, lr_decay_iters=None):
    """ref util.py:8-25. The default policy 'none' is a constant LR via LambdaLR."""
    if lr_policy == 'linear':
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda epoch: max(1.0 - max(0, epoch) / float(n_epochs_decay + 1), 0))
    if lr_policy == 'step':
        return lr_scheduler.StepLR(optimizer, step_size=lr_decay_iters, gamma=0.5)
    if lr_policy == 'plateau':
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.2, threshold=0.01, patience=5)
    if lr_policy == 'cosine':
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=n_epochs, eta_min=0)
    if lr_policy == 'none':
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda x: 1)
    # the reference *returns* (does not raise) the exception object here (util.py:24) — behaviour kept
    return NotImplementedError('learning rate policy [%s] is not implemented', lr_policy)


def get_optimizer(cfg, params):
    """ref util.py:28-39. 'adam' is served by the fused multi-tensor sm_100a kernel (same update rule as
    torch.optim.Adam, same `param_groups` / `state_dict` layout); the other two stay torch's."""
    if cfg['optimizer'] == 'adam':
        from ..optim import FusedAdam

        return FusedAdam(params, lr=cfg['lr'], betas=(cfg['optimizer_beta1'], cfg['optimizer_beta2']))
    if cfg['optimizer'] == 'rmsprop':
        return torch.optim.RMSprop(params, lr=cfg['lr'])
    if cfg['optimizer'] == 'sgd':
        return torch.optim.SGD(params, lr=cfg['lr'])
    return NotImplementedError('optimizer [%s] is not implemented', cfg['optimizer'])


class InputStager:
    """Host -> device staging of one step's inputs on a dedicated copy stream (ref train.py:54-55 moves them on the
    compute stream, where the copy queues behind the previous step's backward pass).

    The returned tensors carry the copy's event as `_splice_ready`; the compute stream waits for it, and LossG lets the
    ViT pass over the step's *target* crops - which depends on nothing but these copies - start right away, i.e. under
    the previous step's backward passes. The `step` scalar stays on the host (Model / LossG accept it either way).

    The device side is a ring of `depth` persistent buffers per input name (views of the right shape are handed out),
    so a step allocates nothing; a buffer is overwritten only after the work that was enqueued between its hand-out and
    the following call has completed (event on the compute stream), however far the host runs ahead."""

    def __init__(self, device=None, depth: int = 12):
        self.device = torch.device('cuda') if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._rings = {}        # input name -> [[buffer, event marking the end of its last use], ...]
        self._handed = []       # slots handed out by the previous call
        self._n = 0

    def __call__(self, batch: dict) -> dict:
        main = torch.cuda.current_stream(self.device)
        if self._handed:        # everything that reads the previous call's buffers has been enqueued by now
            done = torch.cuda.Event()
            done.record(main)
            for slot in self._handed:
                slot[1] = done
            self._handed = []
        out, staged = {}, []
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if k == 'step' or not torch.is_tensor(v) or v.is_cuda:
                    out[k] = v
                    continue
                ring = self._rings.setdefault(k, [[None, None] for _ in range(self.depth)])
                slot = ring[self._n % self.depth]
                n = v.numel()
                if slot[0] is None or slot[0].numel() < n or slot[0].dtype != v.dtype:
                    if slot[0] is not None:
                        slot[0].record_stream(main)      # the outgoing buffer may still be read
                    slot[0] = torch.empty(n + n // 8, dtype=v.dtype, device=self.device)
                if slot[1] is not None:
                    self.stream.wait_event(slot[1])
                d = slot[0][:n].view(v.shape)
                d.copy_(v, non_blocking=True)
                out[k] = d
                staged.append(d)
                self._handed.append(slot)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        main.wait_event(ev)
        for d in staged:
            d._splice_ready = ev
        self._n += 1
        return out


class AsyncScalarLog:
    """Non-blocking stand-in for the per-step `loss_G.item()` of the progress line (ref train.py:67).

    `.item()` drains the stream every step, which leaves the GPU idle while the host enqueues the backward pass. Here
    the scalar is copied into a pinned ring buffer behind an event; `latest()` returns the newest value whose copy has
    landed (at most `depth` steps old - `push` waits for the oldest copy before reusing its slot, so the host never
    runs more than `depth` steps ahead; a deep ring absorbs host-side jitter). `flush()` drains everything (end of the loop)."""

    def __init__(self, depth: int = 8):
        self.depth = depth
        self._host = torch.empty(depth, dtype=torch.float32, pin_memory=True)
        self._events = [torch.cuda.Event() for _ in range(depth)]
        self._busy = [False] * depth
        self._n = 0
        self._newest_done = -1          # sequence number of the newest landed value
        self._seq = [-1] * depth
        self.value = float('nan')

    def _poll(self, block_slot=None):
        for slot in range(self.depth):
            if not self._busy[slot]:
                continue
            if slot == block_slot:
                self._events[slot].synchronize()
            if slot == block_slot or self._events[slot].query():
                self._busy[slot] = False
                if self._seq[slot] > self._newest_done:
                    self._newest_done = self._seq[slot]
                    self.value = float(self._host[slot])

    def push(self, t: torch.Tensor) -> None:
        slot = self._n % self.depth
        if self._busy[slot]:
            self._poll(block_slot=slot)
        self._host[slot:slot + 1].copy_(t.detach().reshape(1), non_blocking=True)
        self._events[slot].record()
        self._busy[slot], self._seq[slot] = True, self._n
        self._n += 1

    def latest(self) -> float:
        self._poll()
        return self.value

    def flush(self) -> float:
        for slot in range(self.depth):
            if self._busy[slot]:
                self._poll(block_slot=slot)
        return self.value


class AsyncImageLog:
    """Deferred `save_result(output[0], dataroot)` + `callback(output[0])` of the image-logging branch (ref train.py:70-76).

    The reference reads the image back right away, which drains the whole stream (the host runs a step or two ahead of
    the device) and then encodes the PNG while the GPU idles. Here the image is copied into pinned host memory behind an
    event; `poll()` - called once per step - hands every image whose copy has landed to ONE worker thread that encodes
    and writes the PNG (150 ms for a 1200x900 image: longer than seven optimisation steps), and fires the callback - on
    the calling thread, like the reference - for every image whose PNG is on disk. Same files, same callback arguments
    in the same order (the tensor handed over lives on the host), a few steps late; `flush()` at the end of the loop
    delivers whatever is still pending."""

    def __init__(self, dataroot, callback=None):
        import queue
        import threading

        self.dataroot, self.callback = dataroot, callback
        self._pending = []          # (event, pinned image), oldest first: copy in flight
        self._encoding = []         # [image, done event, error]: handed to the worker, oldest first
        self._free = []             # pinned buffers ready for reuse (cudaHostAlloc of 13 MB costs milliseconds)
        self._q: "queue.Queue" = queue.Queue()
        self._worker = threading.Thread(target=self._run, name="splice-png", daemon=True)
        self._worker.start()
        self._threading = threading

    def _run(self):
        # torch CPU ops issued from a new thread bring up that thread's own OpenMP team, whose workers spin after every
        # parallel region and starve the (Python-bound) main loop: this thread runs them single-threaded
        try:
            torch.set_num_threads(1)
        except Exception:  # noqa: BLE001
            pass
        while True:
            item = self._q.get()
            if item is None:
                return
            try:
                save_result(item[0], self.dataroot)
            except BaseException as e:  # noqa: BLE001 - surfaced by poll() on the calling thread
                item[2] = e
            item[1].set()

    def push(self, image_t: torch.Tensor) -> None:
        host = None
        for i, buf in enumerate(self._free):
            if buf.shape == image_t.shape and buf.dtype == image_t.dtype:
                host = self._free.pop(i)
                break
        if host is None:
            host = torch.empty(image_t.shape, dtype=image_t.dtype, pin_memory=True)
        host.copy_(image_t.detach(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((ev, host))

    def poll(self, block: bool = False) -> None:
        while self._pending:            # copies that have landed -> the PNG worker, in order
            ev, host = self._pending[0]
            if not block and not ev.query():
                break
            ev.synchronize()
            self._pending.pop(0)
            item = [host, self._threading.Event(), None]
            self._encoding.append(item)
            self._q.put(item)
        while self._encoding:           # PNGs on disk -> callback, in order, on this thread
            host, done, _ = self._encoding[0]
            if not block and not done.is_set():
                break
            done.wait()
            item = self._encoding.pop(0)
            if item[2] is not None:
                raise item[2]
            if self.callback is not None:
                self.callback(host.clone() if len(self._free) < 4 else host)
            if len(self._free) < 4:
                self._free.append(host)

    def flush(self) -> None:
        self.poll(block=True)

    def close(self) -> None:
        self.flush()
        self._q.put(None)
        self._worker.join(timeout=10)


def save_result(image_t, dataroot):
    """Writes <dataroot>/out/output.png (ref util.py:55-59)."""
    from torchvision.transforms import ToPILImage

    out_dir = Path(f"{dataroot}/out")
    out_dir.mkdir(exist_ok=True, parents=True)
    ToPILImage()(image_t).save(f"{out_dir}/output.png")
