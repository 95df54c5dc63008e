"""Shared builders for the parity tests: the same configuration for the CPU oracle (oracle.make_oracle) and the CUDA path
(autorally_b200.scenarios)."""
from autorally_b200.scenarios import (cost_params_for, default_state, make_context, random_network, straight_controls,  # noqa: F401
                                      top_state, warm_controls)
from oracle.oracle import make_oracle  # noqa: F401
