from .config import EngineConfig, config_for, from_oracle_cfg, ARCHS  # noqa: F401
from .core import Engine  # noqa: F401
