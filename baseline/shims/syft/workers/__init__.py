from . import base, virtual, websocket_client, websocket_server  # noqa: F401
