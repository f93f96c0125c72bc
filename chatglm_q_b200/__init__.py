"""chatglm_q_b200 — B200-native (sm_100a) int4g32 / int8 weight-only dequant-matmul path of
K024/chatglm-q: hand-written CUDA kernels behind the reference's QLinear operator surface.

    from chatglm_q_b200 import ops                       # kernel-level (seam S1) functions
    from chatglm_q_b200.int4 import DynamicQuantizeLinear, QEmbedding      # seam S2
    from chatglm_q_b200.install import install           # rebind an imported chatglm_q package

Importing the package does not load the CUDA library; the first op call does, and raises if
`libcgq.so` has not been built (`python -m chatglm_q_b200.build`).  There is no fallback path.
"""
__version__ = "0.1.0"
