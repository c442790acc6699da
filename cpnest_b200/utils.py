"""Logging helpers with the reference's verbosity semantics (cpnest/utils.py:1-100): verbose
0..3 map to CRITICAL / WARNING / INFO / DEBUG on the 'cpnest' logger tree; `LogFile` attaches a
file handler to cpnest.log for the duration of a `with` block."""
import logging

LEVELS = ["CRITICAL", "WARNING", "INFO", "DEBUG"]
FORMATTER = logging.Formatter("%(asctime)s - %(name)-38s: %(message)s", datefmt="%Y-%m-%d, %H:%M:%S")


def _level(verbose):
    return LEVELS[max(0, min(int(verbose), len(LEVELS) - 1))]


class _Handler(object):
    def set_verbosity(self, verbose):
        self.setLevel(_level(verbose))

    def get_verbosity(self):
        return LEVELS.index(logging.getLevelName(self.level))


class StreamHandler(_Handler, logging.StreamHandler):
    def __init__(self, verbose=0, **kwargs):
        logging.StreamHandler.__init__(self, **kwargs)
        self.set_verbosity(verbose)
        self.setFormatter(FORMATTER)


class FileHandler(_Handler, logging.FileHandler):
    def __init__(self, filename, verbose=0, **kwargs):
        logging.FileHandler.__init__(self, filename, **kwargs)
        self.set_verbosity(verbose)
        self.setFormatter(FORMATTER)


class LogFile(object):
    """Context manager adding cpnest.log to the 'cpnest' logger while it is open."""

    def __init__(self, filename, verbose=0, loggername="cpnest"):
        self._filename = filename
        self._verbose = verbose
        self._logger = logging.getLogger(loggername)
        self.handler = None

    def open(self):
        self.handler = FileHandler(self._filename, verbose=self._verbose)
        self._logger.addHandler(self.handler)

    def close(self):
        if self.handler is not None:
            self._logger.removeHandler(self.handler)
            self.handler.close()
            self.handler = None

    def __enter__(self):
        self.open()
        return self

    def __exit__(self, type, value, traceback):
        self.close()
