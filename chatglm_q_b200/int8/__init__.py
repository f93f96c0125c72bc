"""Reference-named aliases for the int8 path (`chatglm_q.int8.qlinear` / `.triton_ops` names)."""
from ..ops import check_input  # noqa: F401
from ..ops import dynamic_quant_matmul as dynamic_quant_matmul_kernel  # noqa: F401
from ..qmodules import W8Embedding as QEmbedding  # noqa: F401
from ..qmodules import W8Linear as DynamicQuantizeLinear  # noqa: F401
from ..qmodules import dynamic_quant_matmul_int8 as dynamic_quant_matmul  # noqa: F401

KERNEL_IMPL = "cgq_b200"
