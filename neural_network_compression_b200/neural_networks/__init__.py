"""Torch counterparts of the reference's Keras models (shapes only matter to the compression path)."""
from .le_net import LeNet5, LeNet300100  # noqa: F401
