"""CPU oracle for the render_mask hot path -- TEST INFRASTRUCTURE, never imported by easyhec_b200."""
from .oracle import *  # noqa: F401,F403
