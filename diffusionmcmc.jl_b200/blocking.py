"""Blocking decorator (/root/reference src/blocking.jl:7-9) and the layout it stands for."""
from .problems import chequerboard_layout  # noqa: F401


class BlockingChequerboardSwitch:
    """Decorator: from here on the updates use the other chequerboard pattern of blocks of
    2*block_half_len inter-observation intervals."""

    def __init__(self, block_half_len):
        self.block_half_len = int(block_half_len)
