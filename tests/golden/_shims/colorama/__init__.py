"""Import stub for colorama (reference utils/logger.py:21)."""


def init(*a, **k):
    return None
