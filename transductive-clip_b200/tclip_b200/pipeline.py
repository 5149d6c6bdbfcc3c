"""Several ``run_task`` batches in flight on one GPU.

The evaluators of the reference run their batches strictly one after the other (``src/eval_zero_shot.py:157-177``:
build the method, ``logs = method.run_task(task_dic=tasks)``, next batch).  On a B200 one ImageNet-shape batch spends
half of its time in a latency-bound tail (a few hundred non-empty clusters iterate 18 x 1000 times; <10 % of the SMs'
issue slots are used), so the whole job — thousands of independent batches — is faster with a few batches in flight:
the tail of one overlaps the throughput-bound head of the next.  Nothing of ``run_task`` changes: every
batch is still one call of the reference-facing method on its own CUDA stream from its own host thread (ctypes and the
CUDA synchronisation calls release the GIL), with its own scratch buffers; results come back in submission order.

    pipe = BatchPipeline(device, streams=3)
    logs = pipe.map(lambda tasks: Method(model=None, device=device, log_file=None, args=args).run_task(task_dic=tasks),
                    batches)
"""
from __future__ import annotations

import sys
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Iterable, List

import torch


_IN_FLIGHT = threading.local()


def batches_in_flight() -> int:
    """How many batches the calling thread's pipeline keeps in flight (1 outside a ``BatchPipeline`` worker).  The method
    classes read it to tell the library (``TCLIP_FLAG_IN_FLIGHT``) that kernels of several batches share the GPU."""
    return getattr(_IN_FLIGHT, "streams", 1)


# While a pipeline with several workers is open the interpreter's thread switch interval is lowered from its default 5 ms:
# a worker that comes back from a CUDA call must get the interpreter quickly to enqueue its batch's next kernels, and with
# batches of ~3 ms (k-means family at RN50 shape) a 5 ms hand-over leaves the GPU idle (measured: 29.8k -> 32.7k tasks/s for
# soft k-means; no effect on the 22 ms EM-Dirichlet batches).  Restored by close().
_SWITCH_INTERVAL = 5e-4


class BatchPipeline:
    def __init__(self, device, streams: int = 3):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tclip_b200 runs on B200 GPUs only: device must be a CUDA device (no CPU fallback)")
        if streams < 1:
            raise ValueError("streams must be >= 1")
        self.streams = int(streams)
        self._local = threading.local()
        self._stream_ids = set()
        self._lock = threading.Lock()
        self._pool = ThreadPoolExecutor(max_workers=self.streams, thread_name_prefix="tclip-batch")
        self._saved_switch_interval = None
        if self.streams > 1 and sys.getswitchinterval() > _SWITCH_INTERVAL:
            self._saved_switch_interval = sys.getswitchinterval()
            sys.setswitchinterval(_SWITCH_INTERVAL)

    def _run(self, fn: Callable, item):
        if getattr(self._local, "stream", None) is None:
            _IN_FLIGHT.streams = self.streams
            torch.cuda.set_device(self.device)
            self._local.stream = torch.cuda.Stream(device=self.device)
            with self._lock:
                self._stream_ids.add(self._local.stream.cuda_stream)
        with torch.cuda.stream(self._local.stream):
            out = fn(item)
            self._local.stream.synchronize()  # the batch is complete (and its scratch reusable) when the call returns
        return out

    def submit(self, fn: Callable, item):
        """Queue one call; returns a ``concurrent.futures.Future`` (for callers that produce their batches one by one, e.g. an
        index sampler that must draw in order on the submitting thread)."""
        return self._pool.submit(self._run, fn, item)

    def map(self, fn: Callable, items: Iterable) -> List:
        """``[fn(item) for item in items]`` with up to ``streams`` calls in flight; results in submission order."""
        futures = [self.submit(fn, it) for it in items]
        return [f.result() for f in futures]

    def close(self):
        self._pool.shutdown(wait=True)
        if self._saved_switch_interval is not None:
            sys.setswitchinterval(self._saved_switch_interval)
            self._saved_switch_interval = None
        from . import ops
        ops.release_workspaces(self._stream_ids)   # the per-stream scratch (1-2 GB each at ImageNet shape)
        self._stream_ids = set()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
