"""`dinov3_jax.data`: the jax-free pieces the training hot path needs live here (collate with the reference's dict
contract, the block-mask generator, a synthetic dataset); the reference's loaders / samplers / datasets /
augmentations "stay" (BASELINE.json north_star) and resolve from a reference checkout placed after this package on
PYTHONPATH (see `dinov3_jax/__init__.py`).  Names are re-exported like the reference's `data/__init__.py:9-13`;
the ones that live in the reference are looked up lazily so that importing this package never requires it."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from .collate import collate_data_and_cast  # noqa: E402,F401
from .masking import MaskingGenerator  # noqa: E402,F401

_LAZY = {  # name -> sub-module of the reference's data package (data/__init__.py:9-13)
    "make_dataset": "loaders", "make_data_loader": "loaders", "SamplerType": "loaders",
    "DataAugmentationDINO": "augmentations",
    "make_classification_eval_transform": "transforms", "make_classification_train_transform": "transforms",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        try:
            mod = importlib.import_module(f"{__name__}.{_LAZY[name]}")
        except ImportError as e:
            raise ImportError(
                f"dinov3_jax.data.{name} lives in the reference's data package (dinov3_jax/data/{_LAZY[name]}.py), which "
                f"this overlay does not replace: put a reference checkout after this package on PYTHONPATH "
                f"(and its own dependencies), or use train.dataset_path=synthetic.  Cause: {e}") from e
        return getattr(mod, name)
    raise AttributeError(name)
