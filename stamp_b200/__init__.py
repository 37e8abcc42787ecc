"""stamp_b200 -- B200 (sm_100a) hot path behind KatherLab/STAMP's plugin interfaces.

Stain normalisation -> ViT tile-feature extraction -> ALiBi Transformer-MIL aggregation as
hand-written CUDA behind a C ABI (``include/stamp_b200.h``); this package is the thin Python
host side that mirrors STAMP's ``Extractor`` / MIL backbone / ``Encoder`` interfaces.
"""

__version__ = "0.1.0"
