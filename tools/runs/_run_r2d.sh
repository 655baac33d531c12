mkdir -p gpurun_out/r2d
tools/ab_phases.sh admm-quantization_b200/lib/libadmmq.so > gpurun_out/r2d/ab.log 2>&1
for db in 0 512 1024 2048 4096; do
  echo "== ADMMQ_DIRECT_BELOW=$db" >> gpurun_out/r2d/direct.log
  for spec in "layer4.1.conv1 36 2" "layer3.1.conv1 7 2" "layer2.1.conv1 2 2" "layer2.1.conv1 2 0"; do
    set -- $spec
    ADMMQ_DIRECT_BELOW=$db python tools/profile_target.py 300 $1 $2 1 $3 >> gpurun_out/r2d/direct.log 2>&1
  done
done
python -m pytest tests -m gpu -q -x -k "not model_top1" > gpurun_out/r2d/pytest.log 2>&1
tail -3 gpurun_out/r2d/pytest.log
cat gpurun_out/r2d/ab.log gpurun_out/r2d/direct.log
