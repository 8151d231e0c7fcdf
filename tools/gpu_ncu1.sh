#!/bin/bash
# One-GPU ncu captures (kernels that do not wait for peers are replayable):  gpurun --timeout 400 -- 'bash tools/gpu_ncu1.sh'
#   * nvl_self_copy_bulk_kernel  - the TMA ring (cp.async.bulk) at 1 GiB: DRAM throughput, instruction mix
#   * nvl_self_copy_kernel       - the thread copy it replaces for big messages
#   * ec_reduce_kernel_t<float,SUM> - ec/cuda one-shot reduce
export PYTHONPATH=$PWD
O=gpurun_out/ncu1
mkdir -p $O
cat > /tmp/ncu_copy.py <<'PY'
import ctypes as C, os, torch
ROOT = os.environ["GRAFT_REPO_ROOT"] if "GRAFT_REPO_ROOT" in os.environ else os.getcwd()
C.CDLL(os.path.join(ROOT, "ucc_b200", "lib", "libucc.so"), mode=C.RTLD_GLOBAL)
nvl = C.CDLL(os.path.join(ROOT, "ucc_b200", "lib", "ucc", "libucc_tl_nvl.so"))
nvl.nvl_launch_self_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
nvl.nvl_launch_self_copy_bulk.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
n = 1 << 28   # 256 MB: twice the L2, keeps the ncu save/restore passes short
src = torch.ones(n // 4, device="cuda"); dst = torch.zeros(n // 4, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    nvl.nvl_launch_self_copy_bulk(dst.data_ptr(), src.data_ptr(), n, 74, None)
    nvl.nvl_launch_self_copy(dst.data_ptr(), src.data_ptr(), n, 256, 512, None)
torch.cuda.synchronize()
assert torch.equal(dst, src)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nvl_self_copy -s 4 -c 2 -o $O/copy_kernels -f python /tmp/ncu_copy.py > $O/ncu_copy.log 2>&1
ncu -i $O/copy_kernels.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    h=rows[0]
    keep=[i for i,c in enumerate(h) if any(k in c for k in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','launch__grid_size','launch__block_size'))]
    for r in rows[:1]+rows[2:]:
        print(' | '.join(r[i] for i in keep))
" > $O/copy_kernels_summary.txt 2>&1
cat $O/copy_kernels_summary.txt | cut -c1-600
tail -3 $O/ncu_copy.log
