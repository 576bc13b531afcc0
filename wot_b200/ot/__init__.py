"""Mirror of the reference's `wot.ot` namespace for the transport-map path (wot/ot/__init__.py:2-6)."""
from .initializer import (initialize_ot_model, parse_configuration, parse_parameter_file,  # noqa: F401
                          parse_per_timepair_configuration, parse_per_timepoint_configuration)
from .optimal_transport import (compute_transport_matrix, last_solve_info, optimal_transport_duality_gap,  # noqa: F401
                                transport_stablev2)
from .ot_model import OTModel  # noqa: F401
from .util import compute_pca  # noqa: F401
