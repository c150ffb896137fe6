"""Thread prefetcher for the scene loop (the step before the path becomes the bottleneck at thousands of
chunks per second).  File reading + parsing (`Dataset.__getitem__`, numpy releases the GIL in its copies and the OS read) runs
`depth` items ahead in a background thread; the consumer -- which issues the H2D copy and the device decode on its own stream --
never waits for the disk unless the reader is slower than the GPU.  Order is preserved; exceptions of the reader are re-raised in
the consumer; the thread ends with the iterator."""
from __future__ import annotations

import queue
import threading

_END = object()


class Prefetcher:
    def __init__(self, iterable, depth=4):
        self._it = iter(iterable)
        self._q = queue.Queue(maxsize=max(1, int(depth)))
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def _run(self):
        try:
            for item in self._it:
                while not self._stop.is_set():
                    try:
                        self._q.put((item, None), timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if self._stop.is_set():
                    return
            self._q.put((_END, None))
        except BaseException as e:  # delivered to the consumer, in order
            self._q.put((_END, e))

    def __iter__(self):
        return self

    def __next__(self):
        item, err = self._q.get()
        if item is _END:
            self._stop.set()
            if err is not None:
                raise err
            raise StopIteration
        return item

    def close(self):
        self._stop.set()
