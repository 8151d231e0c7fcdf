"""Parallelism strategies built on the ucc_b200 collectives: data parallel (bucketed gradient
allreduce overlapped with backward), tensor parallel linear layers, expert-parallel MoE dispatch/combine, ZeRO-1 sharded optimizer step (reduce_scatter + allgather), Ulysses all-to-all and
ring-attention hops (alltoall / neighbour p2p), pipeline stages (GPipe / 1F1B over send / recv)."""
from .ddp import DistributedDataParallel  # noqa: F401
from .tensor_parallel import ColumnParallelLinear, RowParallelLinear  # noqa: F401
from .moe import moe_dispatch, moe_combine  # noqa: F401
from .zero import ZeroRedundancyTrainer, FullyShardedModule  # noqa: F401
from .sequence import ulysses_all_to_all, ring_pass, ring_attention  # noqa: F401
from .pipeline import PipelineStage  # noqa: F401
