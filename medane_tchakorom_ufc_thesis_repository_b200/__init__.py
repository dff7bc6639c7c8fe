"""msplit-b200: B200-native multisplitting solve path (CSR/ELL SpMV, GMRES Arnoldi orthogonalisation,
boundary exchange, TSQR minimisation) behind the reference's operator surface.  See DESIGN.md."""
from ._lib import LIB_PATH, MsplitError  # noqa: F401

__all__ = ["LIB_PATH", "MsplitError"]
