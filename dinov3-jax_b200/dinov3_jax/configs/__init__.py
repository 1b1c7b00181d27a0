from .config import Cfg, DinoV3SetupArgs, apply_scaling_rules_to_cfg, get_cfg_from_args, get_default_config, setup_config  # noqa: F401
