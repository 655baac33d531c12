mkdir -p gpurun_out/r2h
for bn in 0 64 80 96 112 128; do
  echo "== ADMMQ_FOLD_BN=$bn" >> gpurun_out/r2h/mttkrp.log
  if [ $bn = 0 ]; then python tools/time_mttkrp.py 2>&1 | head -3 >> gpurun_out/r2h/mttkrp.log; else ADMMQ_FOLD_BN=$bn python tools/time_mttkrp.py 2>&1 | head -3 >> gpurun_out/r2h/mttkrp.log; fi
done
cat gpurun_out/r2h/mttkrp.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core_mttkrp or contractions or outer_loop_with_tensor" 2>&1 | tail -3
python bench.py --full --init parafac-epc --no-e2e > gpurun_out/r2h/full_resnet18_epc.json 2> gpurun_out/r2h/full_resnet18_epc.err; tail -c 1500 gpurun_out/r2h/full_resnet18_epc.json; tail -3 gpurun_out/r2h/full_resnet18_epc.err
