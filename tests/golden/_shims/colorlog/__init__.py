"""Import stub: the reference's logger imports colorlog (utils/logger.py:18); only needed so the package imports."""
import logging


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, log_colors=None, **kw):
        super().__init__(fmt.replace('%(log_color)s', '') if fmt else fmt, datefmt)
