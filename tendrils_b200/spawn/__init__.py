"""Respawn passes -- mirrors of src/spawn/{init,ball,pixels} of the reference."""
from . import ball, init, pixels  # noqa: F401
from .ball import spawnBall  # noqa: F401
from .init import spawner  # noqa: F401
from .pixels import PixelSpawner  # noqa: F401
