mkdir -p gpurun_out
CMD="python tools/filter_probe.py 200000 784 10000 uniform"
timeout 120 $CMD > gpurun_out/s3e_plain.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bmu_filter_kernel -s 2 -c 1 -o gpurun_out/s3e_filter -f $CMD > gpurun_out/s3e_ncu.log 2>&1; echo "filter capture $?"
timeout 300 ncu --set full --clock-control none -k regex:bmu_refine_kernel -s 2 -c 1 -o gpurun_out/s3e_refine -f $CMD > gpurun_out/s3e_ncu2.log 2>&1; echo "refine capture $?"
CMD2="python tools/filter_train_probe.py 200000 uniform 100"
EPOCHS=45 timeout 200 $CMD2 > gpurun_out/s3e_train_plain.log 2>&1
EPOCHS=45 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/s3e_launches_train.csv $CMD2 > gpurun_out/s3e_train_ncu.log 2>&1; echo "launch list $?"
tail -3 gpurun_out/s3e_plain.log
