"""Control plane: event grammar, pub/sub bus, temporal window, selection, coordinator."""
from .arguments import Arguments  # noqa: F401
from .event_parser import EventParser, Event, STATES, valid_iot_ip_address, format_event  # noqa: F401
from .bus import BusClient, InProcessBroker, TcpBroker, Message, default_broker, reset_default_broker  # noqa: F401
from .window import TemporalWindow, FakeClock, ThreadingTimerFactory  # noqa: F401
from .selection import SelectionPolicy, encrypted_policy  # noqa: F401
