"""Auto-import every method sub-package so that its plugin registers itself
(mirrors ref/src/quantool/methods/__init__.py:9-14: import failures are logged, not fatal)."""
import importlib
import pkgutil

from ..core import LoggerFactory

logger = LoggerFactory.get_logger(__name__)

for _finder, _name, _ispkg in pkgutil.iter_modules(__path__):
    try:
        importlib.import_module(f"{__name__}.{_name}")
        logger.info(f"Imported module: {_name}")
    except Exception as e:  # pragma: no cover
        logger.error(f"Failed to import module {_name}: {e}")
