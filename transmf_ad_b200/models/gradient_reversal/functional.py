"""Gradient-reversal layer (reference models/gradient_reversal/functional.py:4-19): identity in forward,
``-alpha * grad`` in backward, computed by the ``tmf_scale`` kernel."""
from transmf_ad_b200.functional import GradientReversalFunction as GradientReversal

revgrad = GradientReversal.apply
