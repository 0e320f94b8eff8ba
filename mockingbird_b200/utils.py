"""Small host-side helpers mirrored from the reference's utils (utils/util.py:50-61)."""
from __future__ import annotations


class AttrDict(dict):
    """dict with attribute access (utils/util.py:50-53); GAN configs are loaded into this."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """'same' padding of a dilated conv (utils/util.py:60-61)."""
    return int((kernel_size * dilation - dilation) / 2)
