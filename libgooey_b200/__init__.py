"""libgooey_b200 — B200-native batched offline renderer for libgooey's bounce path."""
from ._lib import GooeyError, HostBuffer, LIB_PATH, VoicePatch, lib  # noqa: F401
from . import voices  # noqa: F401
from . import engine  # noqa: F401
from . import bounce  # noqa: F401
