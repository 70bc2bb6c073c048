"""Mirror of the reference's `fmoe` package surface (trainer_3m_fix/fmoe/__init__.py): same class names, constructor
arguments, parameter names and shapes, so 3M-ASR / FastMoE checkpoints load unchanged."""
from .gates import NaiveGate  # noqa: F401
from .layers import FMoE, FMoELinear  # noqa: F401
from .transformer import FMoETransformerMLP  # noqa: F401
