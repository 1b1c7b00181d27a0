from .ssl_meta_arch import SSLMetaArch  # noqa: F401
