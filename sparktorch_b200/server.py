"""Parameter server for ``mode='hogwild'`` (CPU / multi-node transport).

API and HTTP routes follow the reference (sparktorch/server.py:35-151; wire
table SURVEY.md §2.4 X8-X11):

    GET  /            -> "sparktorch"
    GET  /parameters  -> dill(state_dict)                (legacy payload)
    POST /update      <- dill([grad per parameter])      (legacy payload)
    POST /losses      <- {"loss": x} -> {"stop": bool}

plus a binary fast path used by our own workers – the flat fp32 arena goes over
the wire as raw bytes, no pickling:

    GET  /flat/parameters -> version(8B LE) + float32[numel]
    POST /flat/update     <- float32[numel]

Differences: the server is a ``ThreadingHTTPServer`` on a daemon thread of the
driver (no fork, no Flask), the optimizer is the fused flat update
(:func:`apply_update_`), reads take the *read* side of the RWLock and the
early-stop window is guarded by a mutex (quirks Q11/Q12).  On a single B200 node
the GPU-resident owner-sharded path (parallel/hogwild_p2p.py) replaces this
transport entirely.
"""
from __future__ import annotations

import json
import socket
import struct
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Optional

import dill
import numpy as np
import torch

from sparktorch_b200.early_stopper import EarlyStopping
from sparktorch_b200.runtime.flat import FlatArena
from sparktorch_b200.runtime.optim import apply_update_, new_state, parse_optimizer
from sparktorch_b200.rw_lock import RWLock
from sparktorch_b200.util import load_base_torch, load_torch_model


class Server:
    def __init__(self, torch_obj: str, master_url: Optional[str] = None, port: int = 3000,
                 acquire_lock: bool = False, early_stop_patience: int = 1, window_len: int = 4):
        blob, _ = load_base_torch(torch_obj)
        raw = load_torch_model(blob, build_optimizer=False)
        self.model = raw.model
        self.criterion = raw.criterion
        self.model.train()
        self.arena = FlatArena(self.model, torch.device("cpu"))
        self.spec = parse_optimizer(raw.optimizer, raw.optimizer_params)
        self.user_optimizer = None
        if self.spec.kind == "other":
            op = raw.optimizer_params
            self.user_optimizer = raw.optimizer(self.arena.params, **op) if op else raw.optimizer(self.arena.params)
        self.state = new_state(self.spec, self.arena.flat_p) if self.spec.kind != "other" else {}
        self.steps = 0
        self.version = 0
        self.master_url = master_url
        self.port = port
        self.acquire_lock = acquire_lock
        self.window_len = window_len
        self.loss_window = []
        self.should_stop = False
        self.error_count = 0
        self.early_stop_patience = early_stop_patience
        self._lock = RWLock()
        self._loss_mutex = threading.Lock()
        self._es = EarlyStopping(patience=max(early_stop_patience, 1))
        self._httpd: Optional[ThreadingHTTPServer] = None
        self._thread: Optional[threading.Thread] = None

    # ------------------------------------------------------------- lifecycle
    @staticmethod
    def determine_master(port: int) -> str:
        try:
            return socket.gethostbyname(socket.gethostname()) + ":" + str(port)
        except OSError:
            return "localhost:" + str(port)

    def start_server(self) -> None:
        handler = _make_handler(self)
        ThreadingHTTPServer.daemon_threads = True
        ThreadingHTTPServer.allow_reuse_address = True
        # every hogwild worker pulls / pushes once per iteration, all at about the same time: socketserver's default listen backlog
        # of 5 overflows with more than a handful of partitions (dropped handshakes = multi-second stalls, see
        # profiles/hang_8gpu_startup.md for what a backlog of 1 did to the executor start-up)
        ThreadingHTTPServer.request_queue_size = 256
        self._httpd = ThreadingHTTPServer(("0.0.0.0", self.port), handler)
        self.port = self._httpd.server_address[1]
        self._thread = threading.Thread(target=self._httpd.serve_forever, name="sparktorch-ps", daemon=True)
        self._thread.start()
        if not self.master_url or self.master_url.endswith(":0"):
            self.master_url = "127.0.0.1:" + str(self.port)

    def stop_server(self) -> None:
        if self._httpd is not None:
            self._httpd.shutdown()
            self._httpd.server_close()
            self._httpd = None
        if self._thread is not None:
            self._thread.join(timeout=5)
            self._thread = None

    # ------------------------------------------------------------- operations
    def flat_parameters(self) -> bytes:
        if self.acquire_lock:
            self._lock.acquire_read()
        try:
            return struct.pack("<q", self.version) + self.arena.flat_p.numpy().tobytes()
        finally:
            if self.acquire_lock:
                self._lock.release()

    def state_dict_blob(self) -> bytes:
        if self.acquire_lock:
            self._lock.acquire_read()
        try:
            return dill.dumps({k: v.clone() for k, v in self.model.state_dict().items()})
        finally:
            if self.acquire_lock:
                self._lock.release()

    def apply_flat(self, grad: torch.Tensor) -> None:
        if self.acquire_lock:
            self._lock.acquire_write()
        try:
            self.steps += 1
            if self.spec.kind == "other":
                self.arena.grads.copy_(grad)
                self.user_optimizer.step()
            else:
                with torch.no_grad():
                    apply_update_(self.spec, self.arena.flat_p, grad, self.state, self.steps)
            self.version += 1
        except Exception as e:  # tolerate a few malformed pushes like the reference (server.py:141-144)
            self.error_count += 1
            if self.error_count > 10:
                raise RuntimeError(f"Max Errors {e}")
        finally:
            if self.acquire_lock:
                self._lock.release()

    def apply_param_list(self, grads) -> None:
        flat = torch.zeros(self.arena.numel, dtype=torch.float32)
        for g, o, p in zip(grads, self.arena.offsets, self.arena.params):
            if g is not None:
                flat[o:o + p.numel()] = g.detach().reshape(-1).to(torch.float32)
        self.apply_flat(flat)

    def process_loss(self, loss: float) -> bool:
        with self._loss_mutex:
            if self.should_stop:
                return True
            self.loss_window.append(loss)
            if len(self.loss_window) > self.window_len:
                mean = sum(self.loss_window) / len(self.loss_window)
                self.loss_window = []
                if self._es.step(mean):
                    self.should_stop = True
                    return True
            return False

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()}


def _make_handler(ps: Server):
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def _reply(self, payload: bytes, ctype: str = "application/octet-stream", code: int = 200):
            self.send_response(code)
            self.send_header("Content-Type", ctype)
            self.send_header("Content-Length", str(len(payload)))
            self.end_headers()
            self.wfile.write(payload)

        def _body(self) -> bytes:
            n = int(self.headers.get("Content-Length") or 0)
            return self.rfile.read(n) if n else b""

        def do_GET(self):
            if self.path == "/":
                self._reply(b"sparktorch", "text/html; charset=utf-8")
            elif self.path == "/parameters":
                self._reply(ps.state_dict_blob())
            elif self.path == "/flat/parameters":
                self._reply(ps.flat_parameters())
            else:
                self._reply(b"not found", "text/plain", 404)

        def do_POST(self):
            body = self._body()
            if self.path == "/update":
                try:
                    ps.apply_param_list(dill.loads(body))
                except Exception:
                    ps.error_count += 1
                self._reply(b"completed", "text/plain")
            elif self.path == "/flat/update":
                g = torch.from_numpy(np.frombuffer(body, dtype=np.float32).copy())
                ps.apply_flat(g[: ps.arena.numel])
                self._reply(b"completed", "text/plain")
            elif self.path == "/losses":
                stop = ps.process_loss(float(json.loads(body.decode())["loss"]))
                self._reply(json.dumps({"stop": bool(stop)}).encode(), "application/json")
            else:
                self._reply(b"not found", "text/plain", 404)

        def log_message(self, *_a):
            pass

    return Handler
