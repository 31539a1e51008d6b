"""Background feed for the per-step samples of SingleImageDataset.

The reference draws one sample per iteration inside the loop (train.py:53, data/Dataset.py:62-70): PIL colour jitter /
blur / crops on the host - 6 ms per step on a 224 px pair, ~100 ms on the shipped 1200x900 pair - which is longer than
the whole GPU step here. `PrefetchedSamples` moves those calls to one worker thread that runs `depth` samples ahead.

The input stream is unchanged: the worker calls `dataset[0]` in order from a single thread, so numpy's and torch's CPU
RNGs are consumed exactly as in the reference loop (nothing else draws from them once the models are built), and the
`step` tensor - which the dataset mutates in place (Dataset.py:57,63) - is snapshotted per sample. Tensors are pinned
when CUDA is available so that the host -> device copy of InputStager is asynchronous.
"""
from __future__ import annotations

import queue
import threading

import torch


class PrefetchedSamples:
    def __init__(self, dataset, n_samples: int, depth: int = 4, pin: bool | None = None):
        self.dataset = dataset
        self.n = n_samples
        self.pin = torch.cuda.is_available() if pin is None else pin
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

    # pass-through used by the image-logging branch of the loop (train.py:72-73)
    def get_A(self):
        return self.dataset.get_A()
