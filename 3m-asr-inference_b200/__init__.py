"""B200-native fast_moe expert layer (gate -> dispatch -> 32-expert FFN -> combine + residual) of 3M-ASR.

Host side of the drop-in: Python/PyTorch mirrors of the reference's operator surface, all of them thin callers of the
C ABI in include/b200moe.h (libb200moe.so, hand-written sm_100a CUDA).  No Triton, no TensorRT, no CPU fallback.

  ops                                  stage-level wrappers (gate / dispatch / expert_ffn / combine / moe_layer)
  fmoe.{layers,gates,transformer,functions}   mirror of trainer_3m_fix/fmoe (FMoELinear, FMoE, NaiveGate, FMoETransformerMLP, ...)
  layer                                mirror of trainer_3m_fix/layer/positionwise_feed_forward.py (LocalFmoeCatEmbedFeedForward)
  plugin                               mirror of the TensorRT plugins' creator / enqueue / serialize surface
  ep                                   expert parallelism over torch.distributed (NCCL all-to-all over NVLink)
  synth                                seeded synthetic inputs and reference-style random init
"""
from . import _lib, synth  # noqa: F401

__version__ = "0.1.0"
