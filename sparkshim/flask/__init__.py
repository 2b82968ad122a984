"""Minimal Flask look-alike on ``http.server`` (this image has no flask).

Covers what a parameter-server style app needs: ``Flask(name)``,
``@app.route(path, methods=[...])``, the ``request`` proxy (``.data``,
``.json``, ``.method``) and ``app.run(host, port, threaded=True)``.  View
functions may return ``str``, ``bytes`` or a ``dict`` (sent as JSON)."""
from __future__ import annotations

import json
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

__version__ = "0.0+sparkshim"
_ctx = threading.local()


class _Request:
    @property
    def data(self) -> bytes:
        return _ctx.body

    @property
    def json(self):
        return json.loads(_ctx.body.decode() or "null")

    def get_json(self, *_a, **_k):
        return self.json

    @property
    def method(self) -> str:
        return _ctx.method

    @property
    def path(self) -> str:
        return _ctx.path


request = _Request()


def jsonify(obj):
    return obj


class Flask:
    def __init__(self, import_name: str = "app"):
        self.import_name = import_name
        self._routes = {}
        self._server = None

    def route(self, path: str, methods=("GET",)):
        def deco(fn):
            for m in methods:
                self._routes[(m.upper(), path)] = fn
            return fn

        return deco

    def _dispatch(self, method: str, path: str, body: bytes):
        fn = self._routes.get((method, path.split("?")[0]))
        if fn is None:
            return 404, b"not found", "text/plain"
        _ctx.body, _ctx.method, _ctx.path = body, method, path
        try:
            out = fn()
        except Exception as e:  # surfaced as a 500 like flask does
            return 500, repr(e).encode(), "text/plain"
        if isinstance(out, (dict, list)):
            return 200, json.dumps(out).encode(), "application/json"
        if isinstance(out, str):
            return 200, out.encode(), "text/html; charset=utf-8"
        return 200, bytes(out), "application/octet-stream"

    def run(self, host: str = "127.0.0.1", port: int = 5000, debug=None, use_reloader=False, threaded=True, **_k):
        app = self

        class Handler(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def _serve(self, method):
                n = int(self.headers.get("Content-Length") or 0)
                body = self.rfile.read(n) if n else b""
                code, payload, ctype = app._dispatch(method, self.path, body)
                self.send_response(code)
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(payload)))
                self.end_headers()
                self.wfile.write(payload)

            def do_GET(self):
                self._serve("GET")

            def do_POST(self):
                self._serve("POST")

            def log_message(self, *_a):
                pass

        ThreadingHTTPServer.daemon_threads = True
        self._server = ThreadingHTTPServer((host, port), Handler)
        self._server.serve_forever()
