"""Parallelism strategies built on the ucc_b200 collectives: data parallel (bucketed gradient
allreduce overlapped with backward), tensor parallel linear layers, expert-parallel MoE dispatch/combine."""
from .ddp import DistributedDataParallel  # noqa: F401
from .tensor_parallel import ColumnParallelLinear, RowParallelLinear  # noqa: F401
from .moe import moe_dispatch, moe_combine  # noqa: F401
