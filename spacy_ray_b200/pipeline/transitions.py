"""Re-export: the transition systems live next to the model that drives them."""
from ..models.transitions import *  # noqa: F401,F403
from ..models.transitions import (  # noqa: F401
    ArcEagerSystem, ArcState, BiluoState, BiluoSystem, arc_action, arc_decode, biluo_action,
    biluo_actions_to_spans, biluo_decode, is_projective, spans_to_biluo_actions,
)
