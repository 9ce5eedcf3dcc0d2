"""Mirror of the reference's `neural_network_compression.common` package for the compression hot path."""
from . import utility  # noqa: F401
