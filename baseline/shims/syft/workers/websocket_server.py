from .base import BaseWorker


class WebsocketServerWorker(BaseWorker):
    """Imported by the reference's remote_worker.py; the stand-in covers the local path only."""

    def __init__(self, hook=None, host=None, port=None, id="", **kw) -> None:  # noqa: A002
        raise NotImplementedError("stand-in syft: websocket workers are not implemented (local VirtualWorker path only)")
