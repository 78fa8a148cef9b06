"""Library exception for "numerical precision limit reached" conditions.

Same name and role as bayesiancoresets/util/errors.py: raised by the solvers when a greedy step
degenerates (e.g. GIGA's geodesic direction vanishing), caught by the build loops, which retry once
and then latch `reached_numeric_limit`.
"""


class NumericalPrecisionError(Exception):
    """A step could not be taken without losing numerical precision."""
