"""B200-native multi-agent predictive rollout (CfManager / CfAgent hot path of
riddhiman13/predictive-multi-agent-framework). The compute lives in csrc/ (hand-written sm_100a
CUDA behind the C ABI in include/pmaf.h); this package is the host-side mirror used by tests and
bench.py. See DESIGN.md."""
from . import scenarios, loop  # noqa: F401

__all__ = ["scenarios", "loop"]
