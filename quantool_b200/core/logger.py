"""Per-class logger factory.

The reference configures loguru with a rotating file sink created on first import
(ref/src/quantool/core/helpers/logger_factory.py:83-99).  Logging is out of scope for
the hot path; this keeps the same call shape
(``LoggerFactory.get_logger(name)``) on the stdlib logger and writes no files.
"""
import logging
import threading

_lock = threading.Lock()
_configured = False


class LoggerFactory:
    @staticmethod
    def get_logger(name: str) -> logging.Logger:
        global _configured
        with _lock:
            if not _configured:
                root = logging.getLogger("quantool_b200")
                if not root.handlers:
                    h = logging.StreamHandler()
                    h.setFormatter(logging.Formatter(
                        "%(asctime)s %(levelname)-8s [%(name)s] %(message)s"))
                    root.addHandler(h)
                    root.setLevel(logging.WARNING)
                    root.propagate = False
                _configured = True
        return logging.getLogger(f"quantool_b200.{name}")
