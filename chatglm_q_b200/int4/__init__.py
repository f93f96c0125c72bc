"""Reference-named aliases for the int4g32 path (`chatglm_q.int4.qlinear` / `.triton_ops` names)."""
from ..ops import check_input, dynamic_quant_matmul_s4, unpack_int4  # noqa: F401
from ..qmodules import W4Embedding as QEmbedding  # noqa: F401
from ..qmodules import W4Linear as DynamicQuantizeLinear  # noqa: F401
from ..qmodules import dynamic_quant_matmul_int4 as dynamic_quant_matmul  # noqa: F401

DEFAULT_GROUP_SIZE = 32
KERNEL_IMPL = "cgq_b200"
