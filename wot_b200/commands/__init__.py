"""`wot optimal_transport` with the reference's flag set (reference: wot/commands/optimal_transport.py:12-30,
wot/commands/util.py:146-237).  The other eleven wot sub-commands are consumers of transport maps and stay
with the reference package."""
from .optimal_transport import add_ot_parameters_arguments, create_parser, initialize_ot_model_from_args, main  # noqa: F401
