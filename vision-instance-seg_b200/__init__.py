"""B200-native multi-scale deformable attention (the MSDeformAttn hot path of the Swin/R50 +
Mask2Former/MaskDINO stack driven by Wlsghdh/VISION-Instance-Seg's training scripts).

Import name: ``vision_instance_seg_b200`` (the repo-root shim maps it onto this directory).

Public surface — identical to the upstream ``maskdino/modeling/pixel_decoder/ops`` package that the
reference reaches through ``build_model(cfg)`` (/root/reference/training/maskdino/train_full.py:308):

* ``MSDeformAttnFunction.apply(value, spatial_shapes, level_start_index, sampling_locations,
  attention_weights, im2col_step)``
* ``MSDeformAttn(d_model=256, n_levels=4, n_heads=8, n_points=4)``
* ``MultiScaleDeformableAttention.ms_deform_attn_forward / ms_deform_attn_backward`` (the upstream
  extension module's two functions)

All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/msda_b200.h``
(``libmsda_b200.so``).  There is no CPU path: importing works anywhere, calling the operator without
the built library or without a CUDA device raises.
"""
from . import MultiScaleDeformableAttention  # noqa: F401
from ._lib import library_path, load_library  # noqa: F401
from .functions import MSDeformAttnFunction, MSDeformAttnFusedFunction  # noqa: F401
from .modules import (MSDeformAttn, set_fused_encoder_layers, set_fused_preop, share_value_proj,  # noqa: F401
                      unshare_value_proj)



def set_tiled_mode(mode) -> int:
    """Select the kernels of the dense call site (Lq == S, 16-bit values, head dim 32; see include/msda_b200.h
    ``msda_set_tiled_mode``) for this process: 0 / False = direct kernels, 1 / True = tiled forward and backward,
    2 = hybrid backward (direct kernel + sorting kernel for the coarse levels).  Returns the previous mode."""
    return int(load_library().msda_set_tiled_mode(int(mode)))


def set_hybrid_split(adds: int) -> int:
    """Levels that expect more than ``adds`` corner rows per pixel row go to the sorting kernel in mode 2 (see
    include/msda_b200.h ``msda_set_hybrid_split``); returns the previous value."""
    return int(load_library().msda_set_hybrid_split(int(adds)))


def install_as_upstream_extension(name: str = "MultiScaleDeformableAttention"):
    """Register this package's stand-in under the import name of upstream's compiled extension, so that an unmodified
    MaskDINO / Mask2Former / Deformable-DETR checkout — whose ``ops/functions/ms_deform_attn_func.py`` does
    ``import MultiScaleDeformableAttention as MSDA`` — binds to the B200 kernels without building its own extension.

    Call it before the first import of the upstream package, e.g. in the reference's training scripts right before
    ``from maskdino import add_maskdino_config`` (/root/reference/training/maskdino/train_full.py:28)::

        import vision_instance_seg_b200 as b200
        b200.install_as_upstream_extension()

    Returns the module that is now importable as ``name``.  An already imported module of that name is replaced."""
    import sys
    sys.modules[name] = MultiScaleDeformableAttention
    return MultiScaleDeformableAttention


__all__ = ["MSDeformAttn", "MSDeformAttnFunction", "MSDeformAttnFusedFunction", "MultiScaleDeformableAttention",
           "set_fused_preop", "set_fused_encoder_layers", "share_value_proj", "unshare_value_proj", "install_as_upstream_extension", "load_library", "library_path", "set_tiled_mode", "set_hybrid_split"]
