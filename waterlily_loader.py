"""Tiny helper for spawned test workers: `from waterlily_loader import wl` imports the product package."""
import wl_b200 as wl  # noqa: F401
