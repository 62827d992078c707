"""proxsuite.torch (bindings/python/proxsuite/torch/__init__.py): the QP layer on the B200 batch path."""
from .qplayer import QPFunction  # noqa: F401
