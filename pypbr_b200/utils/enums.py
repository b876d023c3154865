"""pypbr_b200.utils.enums — mirrors pypbr/utils/enums.py."""

from enum import Enum


class NormalConvention(Enum):
    OPENGL = 0  # +Y up (normals like [0, 0, 1])
    DIRECTX = 1  # Y axis inverted
