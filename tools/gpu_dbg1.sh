#!/bin/bash
# one-GPU diagnostic: which protocol the active-set p2p test really uses
export PYTHONPATH=$PWD
O=gpurun_out/dbg1; mkdir -p $O
UCC_TL_NVL_LOG_LEVEL=debug timeout 120 python -m pytest tests/test_nvl_gpu.py -m gpu -q -p no:cacheprovider -k "active_set_p2p and 700001" -x -s > $O/p2p.log 2>&1; echo "rc=$?"
grep -c "rendezvous (P2P" $O/p2p.log; grep -c "eager ring (P2P" $O/p2p.log
grep "bytes to\|bytes from" $O/p2p.log | sed 's/^.*TL_NVL *//' | sort | uniq -c | sort -rn | head -20 | cut -c1-200
tail -5 $O/p2p.log | cut -c1-250
