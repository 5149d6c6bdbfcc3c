"""Minimal stand-in for the reference's ``Logger`` (``src/utils.py:171-221``): one ``logging`` logger per module
name with a stream handler and, when a path is given, a file handler; ``del_logger`` drops the handlers
(the reference method classes call it from ``__del__``, ``src/methods/zero_shot/em_dirichlet.py:129-130``).

A method object is built per batch (``src/eval_zero_shot.py:171``) and ``tclip_b200.pipeline`` keeps several alive at
once on different host threads, all sharing the ``logging.Logger`` of their module name.  Handlers are therefore kept
once per (logger name, sink) with a use count: every line is written once per sink however many instances are alive,
and a sink is closed only when its last user calls ``del_logger``."""
from __future__ import annotations

import collections
import logging
import os
import threading

_LOCK = threading.Lock()
_SINKS: dict = {}   # (logger name, sink) -> [handler, use count]; sink = None for the console, else the absolute path
# Releases come from ``__del__`` (the reference's classes close their logger there), i.e. from the garbage collector, which
# can run inside ANY allocation — also inside the critical section below, on the same thread.  A release therefore never
# blocks: it is queued, and applied by whoever holds the lock next (a blocking acquire there would deadlock on itself).
_PENDING: collections.deque = collections.deque()


def _drain_locked():
    while True:
        try:
            logger, key = _PENDING.popleft()
        except IndexError:
            return
        entry = _SINKS.get(key)
        if entry is None:
            continue
        entry[1] -= 1
        if entry[1] <= 0:
            del _SINKS[key]
            logger.removeHandler(entry[0])
            try:
                entry[0].close()
            except Exception:
                pass


def _acquire(logger: logging.Logger, key, make):
    with _LOCK:
        _drain_locked()
        entry = _SINKS.get(key)
        if entry is None:
            handler = make()
            handler.setFormatter(logging.Formatter("[%(name)s]: [%(levelname)s]: %(message)s"))
            logger.addHandler(handler)
            entry = _SINKS[key] = [handler, 0]
        entry[1] += 1


def _release(logger: logging.Logger, key):
    _PENDING.append((logger, key))
    if _LOCK.acquire(blocking=False):
        try:
            _drain_locked()
        finally:
            _LOCK.release()


class Logger:
    def __init__(self, name: str, log_file: str | None = None, level: int = logging.INFO):
        self._logger = logging.getLogger(name)
        self._logger.setLevel(level)
        self._logger.propagate = False
        self._keys = []
        key = (name, None)
        _acquire(self._logger, key, logging.StreamHandler)
        self._keys.append(key)
        if log_file:
            try:
                path = os.path.abspath(log_file)
                d = os.path.dirname(path)
                if d:
                    os.makedirs(d, exist_ok=True)
                key = (name, path)
                _acquire(self._logger, key, lambda: logging.FileHandler(path))
                self._keys.append(key)
            except OSError:
                pass

    def info(self, msg, *a):
        self._logger.info(msg, *a)

    def warning(self, msg, *a):
        self._logger.warning(msg, *a)

    def debug(self, msg, *a):
        self._logger.debug(msg, *a)

    def del_logger(self):
        keys, self._keys = self._keys, []
        for key in keys:
            _release(self._logger, key)
