"""TEST INFRASTRUCTURE ONLY — CPU oracle for the multi-scale deformable attention path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``vision-instance-seg_b200``) never imports this package and has no CPU fallback.

PARITY STATUS: **unpinned against the reference's own tests** — the reference repository
(Wlsghdh/VISION-Instance-Seg) holds no tests, fixtures or golden vectors, and the operator's source
is an un-vendored MaskDINO checkout (``/root/reference/training/maskdino/train_full.py:15-16,28``).
The oracle is instead cross-pinned against an independent implementation that ships in this image
(``transformers`` 5.5 ``MultiScaleDeformableAttention.forward``) through the committed fixtures in
``tests/golden/`` (see ``tests/golden/make_golden.py``) and against a scalar C restatement of the
upstream CUDA kernel semantics (``oracle/msda_oracle.c``).
"""
from .ms_deform_attn_oracle import (  # noqa: F401
    ms_deform_attn_core_pytorch,
    ms_deform_attn_oracle_grads,
    ms_deform_attn_scalar_numpy,
    msdeformattn_preop_pytorch,
    ms_deform_attn_fused_oracle_grads,
)
from .encoder_layer_oracle import add_layernorm_oracle, colsum_oracle, relu_bwd_colsum_oracle  # noqa: F401,E402
