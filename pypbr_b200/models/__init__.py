"""pypbr_b200.models — mirrors pypbr/models/__init__.py."""

from .cooktorrance import BRDFModel, CookTorranceBRDF

__all__ = ["BRDFModel", "CookTorranceBRDF"]
