"""Observation types of the reference's `environments` package (environments/wall_runner.py).  The dm_control simulator
wrapper itself is host-side environment code and out of scope; only the data type that the visual
networks and the visual replay buffer exchange is provided."""
from .wall_runner import MultiObservation  # noqa: F401
