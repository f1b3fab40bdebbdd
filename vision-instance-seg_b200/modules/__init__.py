from .ms_deform_attn import MSDeformAttn, set_fused_preop  # noqa: F401
from .encoder import (MSDeformAttnTransformerEncoder, MSDeformAttnTransformerEncoderLayer,  # noqa: F401
                      MSDeformAttnTransformerEncoderOnly, set_fused_encoder_layers)
from .stacked_value_proj import StackedValueProj, share_value_proj, unshare_value_proj  # noqa: F401
from .decoder import (DeformableTransformerDecoderLayer, TransformerDecoder, build_decoder,  # noqa: F401
                      gen_sineembed_for_position, set_shared_value_proj)
