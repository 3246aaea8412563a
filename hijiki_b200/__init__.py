"""hijiki_b200 — B200-native (sm_100a CUDA) path-tracing hot path of mad-s/hijiki.

The product is the C-ABI library ``hijiki_b200/lib/libhijiki_b200.so`` (``include/hijiki_b200.h``);
this package is the thin host-side mirror of the reference's Rust front-end.
"""
from ._abi import (BLOCK_DTYPE, RAY_DTYPE, EXPORTED_SYMBOLS, HijikiError, HjkParams, HjkStats,
                   HJK_RENDER_ASYNC, HJK_RENDER_EXACT_TIES, HJK_RENDER_KEEP_FEATURES, HJK_RENDER_NO_RECON,
                   LIB_PATH)
from .api import (CompiledScene, Context, DEFAULT_ROOT_SEED, ImageBlockGenerator, Renderer, RenderStats, Scene,
                  comm_unique_id, make_params, split_passes)

__all__ = [
    "BLOCK_DTYPE", "RAY_DTYPE", "EXPORTED_SYMBOLS", "HijikiError", "HjkParams", "HjkStats", "HJK_RENDER_ASYNC",
    "HJK_RENDER_EXACT_TIES", "HJK_RENDER_KEEP_FEATURES", "HJK_RENDER_NO_RECON", "LIB_PATH", "CompiledScene", "Context",
    "DEFAULT_ROOT_SEED", "ImageBlockGenerator", "Renderer", "RenderStats", "Scene", "comm_unique_id",
    "make_params", "split_passes",
]
