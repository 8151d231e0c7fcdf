export PYTHONPATH=$PWD
mkdir -p gpurun_out
for m in direct staged oneshot; do
  k=nvl_reduce_direct_kernel; [ $m = staged ] && k=nvl_reduce_staged_kernel; [ $m = oneshot ] && k=nvl_allreduce_oneshot_kernel
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 --launch-count 1 -f -o gpurun_out/ncu_$m python tools/profile_ncu.py $m > gpurun_out/ncu_$m.log 2>&1
  tail -3 gpurun_out/ncu_$m.log
done
# launch list of a bench run (which of our kernels run, how long)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/profile_ncu.py direct > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
