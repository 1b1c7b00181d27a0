"""`ac_compile_parallelize` (dinov3_jax/fsdp/ac_compile_parallelize.py:20-44): eager sharding of >= 2-D leaves on their
largest divisible axis.  Same policy as `fsdp.utils.shard_params`, with the reference's default threshold."""
from .utils import shard_params


def ac_compile_parallelize(trained_model, inference_only_models=None, config=None, min_shard_size=2 ** 12):
    def only_matrices(tree):
        if isinstance(tree, dict):
            return {k: only_matrices(v) for k, v in tree.items()}
        return tree
    sharded = shard_params(only_matrices(trained_model), "dp", min_param_size=min_shard_size)
    return sharded
