"""Worker handles and the device-side worker runtime.

Replaces PySyft's ``VirtualWorker`` / ``WebsocketClientWorker`` / ``WebsocketServerWorker``
(instantiated at ``federated_coordinator.py:149,167`` and ``remote_worker.py:95``).

* :class:`VirtualWorker` — in-process worker object (local mode, fc.py:139-149).
* :class:`WorkerServer` — device-side runtime (``remote_worker.py`` hosts one): owns the private
  dataset under key ``"training"`` (rw.py:108) and tagged inference tensors (rw.py:102-104) and
  serves ``fit`` / ``search`` / ``predict`` / ``ping`` RPCs.
* :class:`RemoteWorkerClient` — coordinator-side handle with the same verbs.

Wire format: 4-byte big-endian length + msgpack map; tensors travel as raw little-endian fp32
bytes of the **flat arena** (9.6 kB for FFNN instead of ≈36 kB of TorchScript+msgpack per leg,
SURVEY §2.5).  Workers never execute code received from the wire: the model is chosen by name
from the registry and the loss by name — unlike the reference, which ships TorchScript.
"""
from __future__ import annotations

import logging
import socket
import socketserver
import struct
import threading
from typing import Any, Dict, List, Optional, Tuple

import msgpack
import numpy as np
import torch

from ..data import dataset_tensors
from ..fl.trainer import FitConfig, local_fit
from ..fl.evaluate import predict as predict_fn
from ..models import build_model, num_params

log = logging.getLogger(__name__)


class VirtualWorker:
    """In-process worker.  Data arrives via ``federate()`` at training time (reference local mode)
    or can be attached up-front with :meth:`add_dataset`."""

    def __init__(self, worker_id: str, device: Optional[torch.device] = None) -> None:
        self.id = worker_id
        self.device = device
        self.datasets: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.tagged: Dict[str, List[torch.Tensor]] = {}

    def add_dataset(self, dataset, key: str = "training") -> None:
        self.datasets[key] = dataset_tensors(dataset, self.device)

    def load_data(self, tensors: List[torch.Tensor], tag: str = "inference") -> None:
        self.tagged.setdefault(tag, []).extend(tensors)

    def search(self, tag: str) -> List[torch.Tensor]:
        return list(self.tagged.get(tag, []))

    def close(self) -> None:
        pass

    def __repr__(self) -> str:
        return f"<VirtualWorker id:{self.id}>"


# ---------------------------------------------------------------------------------------------
# framing
# ---------------------------------------------------------------------------------------------
def _send_msg(sock: socket.socket, obj: Dict[str, Any]) -> int:
    """Returns the number of bytes put on the wire (frame header included)."""
    data = msgpack.packb(obj, use_bin_type=True)
    sock.sendall(struct.pack(">I", len(data)) + data)
    return 4 + len(data)


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed the connection")
        buf.extend(chunk)
    return bytes(buf)


MAX_FRAME = 1 << 30   # 1 GiB: five times the largest arena in the model zoo (wide MLP, 201.6 MB)


def _recv_msg(sock: socket.socket, counter: Optional[List[int]] = None) -> Dict[str, Any]:
    (n,) = struct.unpack(">I", _recv_exact(sock, 4))
    if n > MAX_FRAME:
        raise ConnectionError(f"frame of {n} bytes exceeds the {MAX_FRAME}-byte limit")
    if counter is not None:
        counter[0] += 4 + n
    return msgpack.unpackb(_recv_exact(sock, n), raw=False)


def _to_bytes(t: torch.Tensor) -> bytes:
    return t.detach().to("cpu", torch.float32).contiguous().numpy().tobytes()


def _from_bytes(b: bytes, device=None) -> torch.Tensor:
    t = torch.from_numpy(np.frombuffer(b, dtype=np.float32).copy())
    return t.to(device) if device is not None else t


# ---------------------------------------------------------------------------------------------
# device-side runtime
# ---------------------------------------------------------------------------------------------
class WorkerServer:
    def __init__(self, worker_id: str, host: str, port: int, device: Optional[torch.device] = None,
                 verbose: bool = False, ssl_context=None) -> None:
        self.id = worker_id
        self.ssl_context = ssl_context          # control/tls.py::server_context(...): the reference's cert_path / key_path stub
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.verbose = verbose
        self.datasets: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.tagged: Dict[str, List[torch.Tensor]] = {}
        self.fits_served = 0
        self._round = 0
        # one model instance + one flat arena per architecture, reused across fits: the trainers behind local_fit
        # (layer-wise MLP, conv net) cache their buffers / CUDA graphs per (model, arena)
        self._models: Dict[str, Tuple[torch.nn.Module, torch.Tensor]] = {}
        self._work_lock = threading.Lock()   # fits / predictions share those buffers: one at a time
        outer = self

        class Handler(socketserver.BaseRequestHandler):
            def setup(self) -> None:
                self.tls_failed = False
                if outer.ssl_context is not None:
                    self.request.settimeout(10.0)
                    try:
                        self.request = outer.ssl_context.wrap_socket(self.request, server_side=True)
                        self.request.settimeout(None)
                    except (OSError, ValueError) as e:
                        log.info("TLS handshake with %s failed: %r", self.client_address, e)
                        self.tls_failed = True

            def handle(self) -> None:
                while not self.tls_failed:
                    try:
                        req = _recv_msg(self.request)
                    except (ConnectionError, OSError, struct.error):
                        return
                    try:
                        resp = outer._dispatch(req)
                    except Exception as e:  # noqa: BLE001 - report to the coordinator, keep serving
                        log.exception("worker RPC failed")
                        resp = {"ok": False, "error": repr(e)}
                    try:
                        _send_msg(self.request, resp)
                    except OSError:
                        return
                    if req.get("op") == "close":
                        return

        class Server(socketserver.ThreadingMixIn, socketserver.TCPServer):
            allow_reuse_address = True
            daemon_threads = True

        self._server = Server((host, port), Handler)
        self.host, self.port = self._server.server_address[:2]
        self._thread: Optional[threading.Thread] = None

    # -- data hosting (rw.py:97-108) -------------------------------------------------------
    def add_dataset(self, dataset, key: str = "training") -> None:
        self.datasets[key] = dataset_tensors(dataset, self.device)

    def load_data(self, tensors: List[torch.Tensor], tag: str = "inference") -> None:
        self.tagged.setdefault(tag, []).extend(t.to(self.device) for t in tensors)

    # -- RPC -----------------------------------------------------------------------------------
    def _dispatch(self, req: Dict[str, Any]) -> Dict[str, Any]:
        op = req.get("op")
        if self.verbose:
            log.info("worker %s <- %s", self.id, op)
        if op == "ping":
            return {"ok": True, "id": self.id}
        if op == "close":
            return {"ok": True}
        if op == "search":
            return {"ok": True, "count": len(self.tagged.get(req["tag"], []))}
        if op == "fit":
            cfg = FitConfig.from_dict(req["config"])
            x, y = self.datasets[req.get("dataset_key", "training")]
            with self._work_lock:
                model, flat, err = self._model_and_arena(cfg.model, req["params"])
                if err:
                    return {"ok": False, "error": err}
                spec = getattr(model, "spec", None)
                if spec is not None and not spec.flatten_input and x.shape[1] != spec.dims[0]:
                    return {"ok": False, "error": f"dataset has {x.shape[1]} features, {cfg.model} wants {spec.dims[0]}"}
                from ..utils.threads import small_model_threads
                with small_model_threads(flat.numel(), flat.device):   # tiny models on a CPU device: no OpenMP fork/join
                    loss, path = local_fit(flat, model, x, y, cfg, round_idx=self._round)
                self._round += 1
                self.fits_served += 1
                reply = {"ok": True, "params": _to_bytes(flat), "loss": float(loss), "n": int(x.shape[0]), "path": path}
                # floating-point buffers the fit updated on this device (BatchNorm running statistics): the flat arena holds
                # named_parameters only, so without them the coordinator would save trained weights next to the INITIAL running
                # stats and every eval-mode use of the checkpoint would be meaningless (FFNN / MLPs have none: nothing is sent)
                bufs = [b.detach().reshape(-1).float() for b in model.buffers() if b.is_floating_point()]
                if bufs:
                    reply["buffers"] = _to_bytes(torch.cat(bufs))
                return reply
        if op == "predict":
            data = self.tagged.get(req.get("tag", "inference"), [])
            if not data:
                return {"ok": True, "pred": [], "count": 0}
            with self._work_lock:
                model, flat, err = self._model_and_arena(req["model"], req["params"])
                if err:
                    return {"ok": False, "error": err}
                spec = getattr(model, "spec", None)
                keep_shape = spec is None                       # conv nets take [C, H, W] samples as they are
                x = torch.stack([d if keep_shape else d.reshape(-1) for d in data]).float()
                pred = predict_fn(model, x, flat)
                return {"ok": True, "pred": pred.reshape(-1).tolist(), "count": len(data)}
        return {"ok": False, "error": f"unknown op {op!r}"}

    def _model_and_arena(self, name: str, params: bytes):
        """(model, arena holding ``params``, error) — instances are created once per architecture."""
        if name not in self._models:
            model = build_model(name).to(self.device)
            self._models[name] = (model, torch.empty(num_params(model), device=self.device))
        model, flat = self._models[name]
        incoming = _from_bytes(params)
        if incoming.numel() != flat.numel():
            return None, None, f"param count {incoming.numel()} != {flat.numel()} for {name}"
        flat.copy_(incoming)
        return model, flat, None

    def start(self, block: bool = True) -> None:
        if block:
            self._server.serve_forever()
        else:
            self._thread = threading.Thread(target=self._server.serve_forever, name=f"worker-{self.id}", daemon=True)
            self._thread.start()

    def stop(self) -> None:
        self._server.shutdown()
        self._server.server_close()


# ---------------------------------------------------------------------------------------------
# coordinator-side handle
# ---------------------------------------------------------------------------------------------
class RemoteWorkerClient:
    def __init__(self, worker_id: str, host: str, port: int, timeout: Optional[float] = 10.0,
                 verbose: bool = False, ssl_context=None) -> None:
        self.id = worker_id
        self.host, self.port = host, port
        self.verbose = verbose
        self._lock = threading.Lock()
        sock = socket.create_connection((host, port), timeout=timeout)
        if ssl_context is not None:
            sock = ssl_context.wrap_socket(sock, server_hostname=host)
        self._sock: Optional[socket.socket] = sock
        self._sock.settimeout(None)
        # application-level traffic of this handle (the paper's §4.3 measures the same two directions with psutil:
        # ~36 kB per worker and round towards the device, ~30.6 kB back, for a 9.6 kB model)
        self.bytes_sent = 0
        self.bytes_received = 0

    def _call(self, req: Dict[str, Any], timeout: Optional[float] = None) -> Dict[str, Any]:
        with self._lock:
            if self._sock is None:
                raise ConnectionError(f"worker {self.id} is closed")
            try:
                self._sock.settimeout(timeout)
                self.bytes_sent += _send_msg(self._sock, req)
                got = [0]
                resp = _recv_msg(self._sock, got)
                self.bytes_received += got[0]
            except (OSError, ConnectionError, struct.error):
                # a timed-out / broken exchange leaves the stream out of step (the late reply would be read as the
                # answer to the next request): the handle is dead from here on
                try:
                    self._sock.close()
                finally:
                    self._sock = None
                raise
        if not resp.get("ok", False):
            raise RuntimeError(f"worker {self.id}: {resp.get('error')}")
        return resp

    def ping(self) -> bool:
        return bool(self._call({"op": "ping"}, timeout=5.0).get("ok"))

    def fit(self, flat: torch.Tensor, cfg: FitConfig, dataset_key: str = "training",
            timeout: Optional[float] = None) -> Tuple[torch.Tensor, float, int]:
        """Broadcast leg + remote local-SGD + gather leg (cf.py:209-211) in one RPC."""
        resp = self._call({"op": "fit", "config": cfg.to_dict(), "params": _to_bytes(flat),
                           "dataset_key": dataset_key}, timeout=timeout)
        self.last_buffers = _from_bytes(resp["buffers"]) if resp.get("buffers") else None     # BatchNorm statistics, if any
        return _from_bytes(resp["params"]), float(resp["loss"]), int(resp["n"])

    def search(self, tag: str) -> int:
        return int(self._call({"op": "search", "tag": tag}).get("count", 0))

    def predict(self, flat: torch.Tensor, model_name: str, tag: str = "inference") -> List[int]:
        return list(self._call({"op": "predict", "params": _to_bytes(flat), "model": model_name, "tag": tag})["pred"])

    def close(self) -> None:
        with self._lock:
            if self._sock is None:
                return
            try:
                _send_msg(self._sock, {"op": "close"})
                _recv_msg(self._sock)
            except (OSError, ConnectionError, struct.error):
                pass
            try:
                self._sock.close()
            finally:
                self._sock = None

    def __repr__(self) -> str:
        return f"<RemoteWorkerClient id:{self.id}>"
