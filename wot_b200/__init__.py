"""wot_b200 -- B200-native transport-map hot path of Waddington-OT (drop-in for wot.ot on that path).

`wot_b200.ot` mirrors the reference's `wot.ot` names for the path:
OTModel, compute_transport_matrix, optimal_transport_duality_gap, transport_stablev2, compute_pca,
parse_configuration, initialize_ot_model.  All arithmetic of the path runs in csrc/libwot_b200.so
(hand-written sm_100a CUDA behind a C ABI, include/wot_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import ot  # noqa: F401,E402
