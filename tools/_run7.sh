python -m pytest tests -m gpu -x -q > gpurun_out/a7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/a7_pytest.log
tail -3 gpurun_out/a7_pytest.log
L=$PWD/streammos_b200/lib
for v in def g4 g8m10; do
  for sp in 0 1 2 4; do
    lib=$L/libstreammos_b200.so; [ $v != def ] && lib=$L/var_$v/libstreammos_b200.so
    echo "$v split=$sp $(SMOS_LIB=$lib SMOS_GATHER_SPLIT=$sp python bench.py --families-only 2>&1 | tail -n 1)" >> gpurun_out/a7_gather.txt
  done
done
cat gpurun_out/a7_gather.txt
