"""Background feed for the per-step samples of SingleImageDataset.

The reference draws one sample per iteration inside the loop (train.py:53, data/Dataset.py:62-70): PIL colour jitter /
blur / crops on the host - 6 ms per step on a 224 px pair, ~100 ms on the shipped 1200x900 pair - which is longer than
the whole GPU step here. `PrefetchedSamples` moves those calls to one worker thread that runs `depth` samples ahead.

The input stream is unchanged: the worker calls `dataset[0]` in order from a single thread, so numpy's and torch's CPU
RNGs are consumed exactly as in the reference loop (nothing else draws from them once the models are built), and the
`step` tensor - which the dataset mutates in place (Dataset.py:57,63) - is snapshotted per sample. Tensors are pinned
when CUDA is available so that the host -> device copy of InputStager is asynchronous.

`mode="process"` runs the same worker in a forked child instead of a thread: PIL's augmentation holds the GIL for most of
its 6 ms (224 px) and the optimisation loop itself is ~2 ms of Python per step, so a worker THREAD serialises with the
loop (measured 127 it/s against 180 for the bare step); a forked child inherits the generator states at the moment the
feed is created - exactly the states the inline loop would draw from - and never touches CUDA (PIL / CPU tensors only;
samples come back through torch.multiprocessing's shared memory and are pinned by the parent). Device-side datasets
(data/device_aug.py) need the thread mode.
"""
from __future__ import annotations

import queue
import threading

import torch


def _child_main(dataset, n, q, stop):
    """Body of the forked worker: no CUDA, single-threaded torch ops (an OpenMP team inherited through fork may hang)."""
    import gc
    import os

    # The child inherits every Python object of the parent, CUDA tensors of earlier runs included. Freeing one of them here
    # (cyclic GC) calls into a CUDA runtime that is not usable after fork ("initialization error" -> std::terminate):
    # freeze what was inherited, collect nothing, and leave through os._exit so that no finaliser runs either.
    gc.disable()
    gc.freeze()
    try:
        torch.set_num_threads(1)
        for _ in range(n):
            if stop.is_set():
                return
            sample = dataset[0]
            out = {k: ((v.clone() if k == 'step' else v.contiguous()) if torch.is_tensor(v) else v) for k, v in sample.items()}
            while not stop.is_set():
                try:
                    q.put(out, timeout=0.1)
                    break
                except queue.Full:
                    continue
        # the tensors travel as file descriptors served by THIS process: stay until the consumer has taken the last one
        stop.wait(timeout=3600)
    except BaseException as e:  # noqa: BLE001 - surfaced to the consumer
        try:
            q.put(RuntimeError(f"prefetch worker failed: {e!r}"))
            stop.wait(timeout=60)
        except BaseException:  # noqa: BLE001
            pass
    finally:
        try:
            q.close()
            q.join_thread()
        except BaseException:  # noqa: BLE001
            pass
        os._exit(0)


class PrefetchedSamples:
    def __init__(self, dataset, n_samples: int, depth: int = 4, pin: bool | None = None, mode: str = "thread"):
        self.dataset = dataset
        self.n = n_samples
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.mode = mode
        if mode == "process":
            import torch.multiprocessing as mp

            ctx = mp.get_context("fork")
            self._q = ctx.Queue(maxsize=max(1, depth))
            self._stop = ctx.Event()
            self._worker = ctx.Process(target=_child_main, args=(dataset, n_samples, self._q, self._stop), daemon=True)
            self._worker.start()
            return
        if mode != "thread":
            raise ValueError(f"prefetch mode {mode!r}: 'thread' or 'process'")
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._stop = threading.Event()
        # torch.default_generator, numpy's global RandomState and python's `random` are process-wide: the worker continues
        # the very streams the main thread seeded
        self._worker = threading.Thread(target=self._run, name="splice-prefetch", daemon=True)
        self._worker.start()

    def _run(self):
        try:
            # this thread's torch CPU ops (ToTensor, crops) single-threaded: a second OpenMP team that spins after every
            # parallel region costs the main loop more than the few hundred microseconds it saves here (OpenMP's thread
            # count is a per-thread setting: the main thread keeps its own)
            try:
                torch.set_num_threads(1)
            except Exception:  # noqa: BLE001
                pass
            for _ in range(self.n):
                if self._stop.is_set():
                    return
                sample = self.dataset[0]
                out = {}
                for k, v in sample.items():
                    if torch.is_tensor(v):
                        v = v.clone() if k == 'step' else v.contiguous()
                        if self.pin and k != 'step' and not v.is_cuda:     # device-side feeds hand over CUDA tensors
                            v = v.pin_memory()
                    out[k] = v
                while not self._stop.is_set():
                    try:
                        self._q.put(out, timeout=0.1)
                        break
                    except queue.Full:
                        continue
        except BaseException as e:  # noqa: BLE001 - surfaced to the consumer
            self._q.put(e)

    def next(self) -> dict:
        item = self._q.get()
        if isinstance(item, BaseException):
            raise item
        if self.mode == "process" and self.pin:     # the child cannot pin (no CUDA there): one host copy per tensor here
            item = {k: (v.pin_memory() if torch.is_tensor(v) and k != 'step' else v) for k, v in item.items()}
        return item

    def __iter__(self):
        for _ in range(self.n):
            yield self.next()

    def close(self):
        self._stop.set()
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass
        self._worker.join(timeout=5)
        if self.mode == "process" and self._worker.is_alive():
            self._worker.terminate()

    # pass-through used by the image-logging branch of the loop (train.py:72-73)
    def get_A(self):
        return self.dataset.get_A()
