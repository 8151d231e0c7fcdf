"""ucc_b200 — B200-native collective communication library with UCC's C API.

The product is the native library (``ucc_b200/lib/libucc.so`` + sm_100a
plugin modules under ``ucc_b200/lib/ucc``).  This package provides

* :mod:`ucc_b200.capi`     – ctypes binding of the C API,
* :mod:`ucc_b200.harness`  – single-process multi-rank job (test harness),
* :mod:`ucc_b200.dist`     – one-process-per-GPU bootstrap on top of torch.distributed,
* :mod:`ucc_b200.ops`      – collectives on torch tensors,
* :mod:`ucc_b200.parallel` – DP/FSDP/TP/SP/EP/PP communication patterns built on them.
"""
__version__ = "0.1.0"
