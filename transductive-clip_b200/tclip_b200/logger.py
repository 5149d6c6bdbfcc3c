"""Minimal stand-in for the reference's ``Logger`` (``src/utils.py:171-221``): one ``logging`` logger per module
name with a stream handler and, when a path is given, a file handler; ``del_logger`` drops the handlers
(the reference method classes call it from ``__del__``, ``src/methods/zero_shot/em_dirichlet.py:129-130``)."""
from __future__ import annotations

import logging
import os


class Logger:
    def __init__(self, name: str, log_file: str | None = None, level: int = logging.INFO):
        self._logger = logging.getLogger(name)
        self._logger.setLevel(level)
        self._logger.propagate = False
        self._handlers = []
        fmt = logging.Formatter("[%(name)s]: [%(levelname)s]: %(message)s")
        if not any(isinstance(h, logging.StreamHandler) and not isinstance(h, logging.FileHandler)
                   for h in self._logger.handlers):
            sh = logging.StreamHandler()
            sh.setFormatter(fmt)
            self._logger.addHandler(sh)
            self._handlers.append(sh)
        if log_file:
            try:
                d = os.path.dirname(log_file)
                if d:
                    os.makedirs(d, exist_ok=True)
                fh = logging.FileHandler(log_file)
                fh.setFormatter(fmt)
                self._logger.addHandler(fh)
                self._handlers.append(fh)
            except OSError:
                pass

    def info(self, msg, *a):
        self._logger.info(msg, *a)

    def warning(self, msg, *a):
        self._logger.warning(msg, *a)

    def debug(self, msg, *a):
        self._logger.debug(msg, *a)

    def del_logger(self):
        for h in self._handlers:
            self._logger.removeHandler(h)
            try:
                h.close()
            except Exception:
                pass
        self._handlers = []
