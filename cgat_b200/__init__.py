"""cgat_b200 — B200-native implementation of CGAT's edge-wise graph-attention hot path.

`from cgat_b200 import CGAtNet` mirrors the reference's `from CGAT import CGAtNet`
(reference CGAT/__init__.py:1); `--version cgat_b200.CGAT` selects it through the reference's own
plugin switch (reference CGAT/lightning_module.py:165-166, 451-454)."""
from .CGAT import CGAtNet  # noqa: F401

__all__ = ["CGAtNet"]
