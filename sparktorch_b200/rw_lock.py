"""Writer-priority reader/writer lock.

Same contract as the reference lock (sparktorch/rw_lock.py:11-66): many readers
XOR one writer, waiting writers block new readers, one ``release()`` for both
kinds.  Implemented on a single ``threading.Condition``; additionally usable as
context managers (``with lock.reading(): ...``)."""
from __future__ import annotations

import threading
from contextlib import contextmanager


class RWLock:
    def __init__(self):
        self._cv = threading.Condition()
        self._readers = 0          # active readers
        self._writer = False       # a writer holds the lock
        self._writers_waiting = 0

    def acquire_read(self) -> None:
        with self._cv:
            while self._writer or self._writers_waiting:
                self._cv.wait()
            self._readers += 1

    def acquire_write(self) -> None:
        with self._cv:
            self._writers_waiting += 1
            try:
                while self._writer or self._readers:
                    self._cv.wait()
            finally:
                self._writers_waiting -= 1
            self._writer = True

    def release(self) -> None:
        with self._cv:
            if self._writer:
                self._writer = False
            elif self._readers:
                self._readers -= 1
            self._cv.notify_all()

    @contextmanager
    def reading(self):
        self.acquire_read()
        try:
            yield
        finally:
            self.release()

    @contextmanager
    def writing(self):
        self.acquire_write()
        try:
            yield
        finally:
            self.release()
