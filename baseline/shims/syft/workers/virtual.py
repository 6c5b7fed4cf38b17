from .base import BaseWorker


class VirtualWorker(BaseWorker):
    """In-process worker (``sy.VirtualWorker(hook, id)``)."""

    def __init__(self, hook=None, id: str = "", **kw) -> None:  # noqa: A002
        super().__init__(hook, id, **kw)
