#!/bin/bash
# one-GPU diagnostic: the active-set p2p tests, three times (the rendezvous hang was intermittent), then the whole GPU gate
export PYTHONPATH=$PWD
O=gpurun_out/dbg1; mkdir -p $O
for i in 1 2 3; do
  UCC_TL_NVL_LOG_LEVEL=debug timeout 100 python -m pytest tests/test_nvl_gpu.py -m gpu -q -p no:cacheprovider -k "active_set_p2p" > $O/p2p_$i.log 2>&1; echo "p2p run $i rc=$? $(tail -1 $O/p2p_$i.log)"
done
grep -h "Error\|ERROR\|^E  " $O/p2p_*.log | head -10 | cut -c1-220
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest.log 2>&1; echo "gate rc=$?"; tail -4 $O/pytest.log | cut -c1-250
